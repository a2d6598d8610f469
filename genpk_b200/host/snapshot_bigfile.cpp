// Minimal bigfile reader: what gen-pk needs from an MP-Gadget snapshot
// (read_fieldize_bigfile.cpp:7-125).  On-disk format (bigfile/src/bigfile.c): a
// snapshot is a directory, a block a sub-directory holding a text `header`
// ("DTYPE: <f8", "NMEMB: 3", "NFILE: 2", then one "%06X: rows : checksum : sysv" line
// per data file, bigfile.c:506-528), `attr-v2` (one "name dtype nmemb HEXBYTES #HUMANE
// [...]" line per attribute, bigfile.c:1452-1513) and raw little-endian row-major data
// files 000000, 000001, ...  Checksums are not verified (the reference never does).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <fstream>
#include <sstream>

#include "snapshot.hpp"

namespace genpk_host {

static int itemsize_of(const std::string &dtype)
{
    if (dtype.size() < 2)
        return 0;
    return atoi(dtype.c_str() + (strchr("<>=|", dtype[0]) ? 2 : 1));
}
static char kind_of(const std::string &dtype) { return dtype.size() < 2 ? 0 : (strchr("<>=|", dtype[0]) ? dtype[1] : dtype[0]); }

static double element_as_double(const unsigned char *p, char kind, int size)
{
    switch (kind) {
    case 'f':
        if (size == 8) { double v; memcpy(&v, p, 8); return v; }
        if (size == 4) { float v; memcpy(&v, p, 4); return v; }
        break;
    case 'i':
        if (size == 8) { int64_t v; memcpy(&v, p, 8); return (double)v; }
        if (size == 4) { int32_t v; memcpy(&v, p, 4); return (double)v; }
        break;
    case 'u':
        if (size == 8) { uint64_t v; memcpy(&v, p, 8); return (double)v; }
        if (size == 4) { uint32_t v; memcpy(&v, p, 4); return (double)v; }
        break;
    }
    return 0.0;
}

BigfileSnapshot::BigfileSnapshot(const std::string &dir) : dir_(dir)
{
    std::ifstream hdr(dir + "/Header/header");
    if (!hdr)
        return;                                   // not a bigfile: is_bigfile() == 0
    ok_ = true;
    std::ifstream at(dir + "/Header/attr-v2");
    std::string line;
    while (std::getline(at, line)) {
        std::istringstream ls(line);
        std::string name, dtype, hex;
        int nmemb = 0;
        if (!(ls >> name >> dtype >> nmemb >> hex))
            continue;
        Attr a;
        a.dtype = dtype;
        a.nmemb = nmemb;
        for (size_t i = 0; i + 1 < hex.size(); i += 2)
            a.raw.push_back((unsigned char)strtol(hex.substr(i, 2).c_str(), nullptr, 16));
        attrs_[name] = a;
    }
}

bool BigfileSnapshot::attr_f8(const std::string &name, double *out, int n) const
{
    auto it = attrs_.find(name);
    if (it == attrs_.end() || it->second.nmemb < n) {
        error_ = "attribute " + name + " missing in Header";
        return false;
    }
    const int sz = itemsize_of(it->second.dtype);
    if ((int)it->second.raw.size() < sz * n)
        return false;
    for (int i = 0; i < n; i++)
        out[i] = element_as_double(it->second.raw.data() + (size_t)i * sz, kind_of(it->second.dtype), sz);
    return true;
}

bool BigfileSnapshot::attr_i8(const std::string &name, int64_t *out, int n) const
{
    auto it = attrs_.find(name);
    if (it == attrs_.end() || it->second.nmemb < n) {
        error_ = "attribute " + name + " missing in Header";
        return false;
    }
    const int sz = itemsize_of(it->second.dtype);
    const char kind = kind_of(it->second.dtype);
    if ((int)it->second.raw.size() < sz * n)
        return false;
    for (int i = 0; i < n; i++) {
        const unsigned char *p = it->second.raw.data() + (size_t)i * sz;
        if ((kind == 'i' || kind == 'u') && sz == 8) {
            memcpy(&out[i], p, 8);
        } else {
            out[i] = (int64_t)element_as_double(p, kind, sz);
        }
    }
    return true;
}

bool BigfileSnapshot::open_block(const std::string &name, BigBlockInfo *info) const
{
    info->dir = dir_ + "/" + name;
    std::ifstream hdr(info->dir + "/header");
    if (!hdr) {
        error_ = "no block " + name + " in " + dir_;
        return false;
    }
    std::string line;
    int nfile = 0;
    info->file_rows.clear();
    info->rows = 0;
    while (std::getline(hdr, line)) {
        if (line.compare(0, 6, "DTYPE:") == 0) {
            std::istringstream(line.substr(6)) >> info->dtype;
        } else if (line.compare(0, 6, "NMEMB:") == 0) {
            info->nmemb = atoi(line.c_str() + 6);
        } else if (line.compare(0, 6, "NFILE:") == 0) {
            nfile = atoi(line.c_str() + 6);
        } else if (line.size() > 7 && line[6] == ':') {
            const int64_t rows = strtoll(line.c_str() + 7, nullptr, 10);
            info->file_rows.push_back(rows);
            info->rows += rows;
        }
    }
    info->itemsize = itemsize_of(info->dtype);
    if ((int)info->file_rows.size() != nfile || info->itemsize == 0 || info->dtype[0] == '>') {
        error_ = "unsupported or corrupt header in block " + name;
        return false;
    }
    return true;
}

// rows [first, first + count) of a block as byte ranges of its files, laid out back to back from dst
static bool plan_rows(const BigBlockInfo &b, int64_t first, int64_t count, size_t rowbytes, void *dst, std::vector<ReadSeg> *segs)
{
    int64_t file_first = 0, done = 0;
    for (size_t f = 0; f < b.file_rows.size() && done < count; f++) {
        const int64_t file_last = file_first + b.file_rows[f];
        const int64_t lo = first + done;
        if (lo < file_last) {
            const int64_t n = (file_last - lo < count - done) ? file_last - lo : count - done;
            char fname[16];
            snprintf(fname, sizeof(fname), "%06X", (unsigned)f);
            ReadSeg g;
            g.path = b.dir + "/" + fname;
            g.offset = (lo - file_first) * (int64_t)rowbytes;
            g.bytes = n * (int64_t)rowbytes;
            g.dst = (char *)dst + (size_t)done * rowbytes;
            segs->push_back(g);
            done += n;
        }
        file_first = file_last;
    }
    return done == count;
}

bool BigfileSnapshot::read_f32(const BigBlockInfo &b, int64_t first, int64_t count, float *dst) const
{
    const char kind = kind_of(b.dtype);
    const size_t rowbytes = (size_t)b.itemsize * b.nmemb;
    const size_t items = (size_t)count * b.nmemb;
    const bool as_stored = kind == 'f' && b.itemsize == 4;
    std::vector<unsigned char> buf;
    if (!as_stored)
        buf.resize((size_t)count * rowbytes);
    std::vector<ReadSeg> segs;
    if (!plan_rows(b, first, count, rowbytes, as_stored ? (void *)dst : (void *)buf.data(), &segs)) {
        error_ = "block " + b.dir + " holds fewer rows than requested";
        return false;
    }
    if (!read_segments(segs, read_threads(), &error_))
        return false;
    if (!as_stored) {
        // through double, then narrowed: positions[i] = ((double*)pos.data)[i] (read_fieldize_bigfile.cpp:93-94)
        for (size_t i = 0; i < items; i++)
            dst[i] = (float)element_as_double(buf.data() + i * b.itemsize, kind, b.itemsize);
    }
    return true;
}

bool BigfileSnapshot::read_f64_raw(const BigBlockInfo &b, int64_t first, int64_t count, double *dst) const
{
    if (b.dtype != "<f8") {
        error_ = "block " + b.dir + " is not <f8";
        return false;
    }
    std::vector<ReadSeg> segs;
    if (!plan_rows(b, first, count, (size_t)8 * b.nmemb, dst, &segs)) {
        error_ = "block " + b.dir + " holds fewer rows than requested";
        return false;
    }
    return read_segments(segs, read_threads(), &error_);
}

}  // namespace genpk_host
