// Gadget format II reader: block map per file + ranged reads.  Behaviour follows
// GadgetReader (gadgetreader.cpp:52-120 file set, :124-282 block scan, :290-316 block
// head, :471-557 GetBlock); little-endian format-II files only (what every BASELINE
// config uses) -- byte-swapped and name-less Gadget-I files are reported as unsupported.
#include <stdio.h>
#include <string.h>

#include "snapshot.hpp"

namespace genpk_host {

static bool read_u32(FILE *fd, uint32_t *v) { return fread(v, 4, 1, fd) == 1; }

bool GadgetSnapshot::scan_file(const std::string &path, GadgetFile *out)
{
    FILE *fd = fopen(path.c_str(), "rb");
    if (!fd)
        return false;
    out->name = path;
    int64_t total_file_part = 0;
    bool have_head = false;
    for (;;) {
        // block head record: [8]["NAME"][datalen+8][8], then [datalen][data][datalen]
        uint32_t head[4];
        if (fread(head, 4, 4, fd) != 4)
            break;
        if (head[0] != 8 || head[3] != 8) {
            if (!have_head)
                error_ = path + ": not a little-endian Gadget format II file (Gadget-I / byte-swapped files are not supported)";
            break;
        }
        char name[5] = {0, 0, 0, 0, 0};
        memcpy(name, &head[1], 4);
        uint32_t reclen = 0;
        if (!read_u32(fd, &reclen) || reclen != head[2] - 8)
            break;
        if (strncmp(name, "HEAD", 4) == 0) {
            uint32_t tail = 0;
            if (reclen != sizeof(GadgetHeader) || fread(&out->header, sizeof(GadgetHeader), 1, fd) != 1 ||
                !read_u32(fd, &tail) || tail != reclen)
                break;
            for (int t = 0; t < N_TYPE; t++)
                total_file_part += out->header.npart[t];
            have_head = true;
            continue;
        }
        GadgetBlock b;
        b.length = reclen;
        // bytes per particle: the heuristics of gadgetreader.cpp:211-232
        if (strncmp(name, "POS ", 4) == 0 || strncmp(name, "VEL ", 4) == 0)
            b.partlen = (b.length == 3 * total_file_part * 8) ? 24 : 12;
        else if (strncmp(name, "ID  ", 4) == 0)
            b.partlen = (b.length == total_file_part * 4) ? 4 : 8;
        else
            b.partlen = (b.length == total_file_part * 8) ? 8 : 4;
        b.start = ftell(fd);
        uint32_t tail = 0;
        if (fseek(fd, b.length, SEEK_CUR) != 0 || !read_u32(fd, &tail) || tail != reclen)
            break;
        std::string key(name);
        while (out->blocks.count(key))       // duplicate names move ahead one (gadgetreader.cpp:268-271)
            key[3]++;
        out->blocks[key] = b;
    }
    fclose(fd);
    return have_head;
}

GadgetSnapshot::GadgetSnapshot(const std::string &base)
{
    std::string first = base;
    FILE *fd = fopen(first.c_str(), "rb");
    if (!fd) {
        first += ".0";
        fd = fopen(first.c_str(), "rb");
    }
    if (!fd) {
        error_ = "could not open " + base + " (.0)";
        return;
    }
    fclose(fd);
    GadgetFile f0;
    if (!scan_file(first, &f0)) {
        if (error_.empty())
            error_ = first + ": no HEAD block";
        return;
    }
    std::string stem = first;
    if (stem.size() > 2 && stem.compare(stem.size() - 2, 2, ".0") == 0)
        stem.erase(stem.size() - 2);
    files_.push_back(f0);
    const int expected = f0.header.num_files;
    if (expected < 1 || expected > 999)
        return;
    for (int i = 1; i < expected; i++) {
        GadgetFile fi;
        if (scan_file(stem + "." + std::to_string(i), &fi) && !fi.blocks.empty())
            files_.push_back(fi);
    }
}

int64_t GadgetSnapshot::npart(int type) const
{
    if (type < 0 || type >= N_TYPE || files_.empty())
        return 0;
    // long word + low word of the totals in file 0 (GSnap::GetNpart, gadgetreader.cpp:396-411);
    // a long word that disagrees with what the files hold is treated as bogus (:109-113)
    int64_t found = 0;
    for (const auto &f : files_)
        found += f.header.npart[type];
    const int64_t lo = files_[0].header.npartTotal[type];
    const int64_t full = ((int64_t)files_[0].header.NallHW[type] << 32) + lo;
    return (full != found && lo == found) ? lo : full;
}

bool GadgetSnapshot::has_block(const std::string &name) const
{
    for (const auto &f : files_)
        if (f.blocks.count(name))
            return true;
    return false;
}

int64_t GadgetSnapshot::block_bytes(const std::string &name) const
{
    int64_t s = 0;
    for (const auto &f : files_) {
        auto it = f.blocks.find(name);
        if (it != f.blocks.end())
            s += it->second.length;
    }
    return s;
}

int64_t GadgetSnapshot::block_parts(const std::string &name) const
{
    int64_t s = 0;
    for (const auto &f : files_) {
        auto it = f.blocks.find(name);
        if (it != f.blocks.end())
            s += it->second.length / it->second.partlen;
    }
    return s;
}

// The pieces of the block that a call wants, file by file: [n_file particles at file offset pos], in file order.
int64_t GadgetSnapshot::get_block(const std::string &name, void *dst, int64_t n_to_read, int64_t start_part,
                                  int skip_type) const
{
    if (!has_block(name))
        return 0;
    const int64_t start_part_in = start_part;
    struct Want {
        const GadgetFile *file;
        int64_t pos;
        uint32_t n_file;
        int partlen;
    };
    std::vector<Want> wants;
    int64_t planned = 0;
    for (const auto &f : files_) {
        auto it = f.blocks.find(name);
        if (it == f.blocks.end())
            continue;
        const GadgetBlock &b = it->second;
        // particles of this file that are wanted: 32-bit unsigned arithmetic, as in the reference --
        // when skip_type names a type the block does not hold this wraps, and the clamp below is
        // what bounds the read
        uint32_t n_file = (uint32_t)(b.length / b.partlen);
        for (int j = 0; j < N_TYPE; j++)
            if (skip_type & (1 << j))
                n_file -= f.header.npart[j];
        int64_t pos = b.start;
        for (int j = 0; j < N_TYPE; j++) {
            if (skip_type & (1 << j))
                pos += (int64_t)f.header.npart[j] * b.partlen;
            else
                break;
        }
        if (start_part > 0) {
            if ((int64_t)n_file <= start_part) {
                start_part -= n_file;
                continue;
            }
            pos += start_part * b.partlen;
            n_file -= (uint32_t)start_part;
            start_part = 0;
        }
        if ((int64_t)n_file > n_to_read - planned)
            n_file = (uint32_t)(n_to_read - planned);
        wants.push_back({&f, pos, n_file, b.partlen});
        planned += n_file;
        if (planned == n_to_read)
            break;
    }
    // all pieces at once (parallel_read.cpp); every piece complete is the normal case
    {
        std::vector<ReadSeg> segs;
        int64_t at = 0;
        for (const Want &w : wants) {
            ReadSeg g;
            g.path = w.file->name;
            g.offset = w.pos;
            g.bytes = (int64_t)w.n_file * w.partlen;
            g.dst = (char *)dst + at * w.partlen;
            segs.push_back(g);
            at += w.n_file;
        }
        if (read_segments(segs, read_threads(), nullptr))
            return planned;
    }
    // a missing file or a short read: exactly what the reference's reader would have delivered
    return get_block_sequential(name, dst, n_to_read, start_part_in, skip_type);
}

// File by file through fread(), as GSnap::GetBlock does (gadgetreader.cpp:471-557): a file that cannot be opened is
// skipped and what fread() returned is what counts.
int64_t GadgetSnapshot::get_block_sequential(const std::string &name, void *dst, int64_t n_to_read, int64_t start_part,
                                  int skip_type) const
{
    int64_t n_read = 0;
    if (!has_block(name))
        return 0;
    for (const auto &f : files_) {
        auto it = f.blocks.find(name);
        if (it == f.blocks.end())
            continue;
        const GadgetBlock &b = it->second;
        // particles of this file that are wanted: 32-bit unsigned arithmetic, as in the reference --
        // when skip_type names a type the block does not hold this wraps, and the clamp below is
        // what bounds the read
        uint32_t n_file = (uint32_t)(b.length / b.partlen);
        for (int j = 0; j < N_TYPE; j++)
            if (skip_type & (1 << j))
                n_file -= f.header.npart[j];
        int64_t pos = b.start;
        for (int j = 0; j < N_TYPE; j++) {
            if (skip_type & (1 << j))
                pos += (int64_t)f.header.npart[j] * b.partlen;
            else
                break;
        }
        if (start_part > 0) {
            if ((int64_t)n_file <= start_part) {
                start_part -= n_file;
                continue;
            }
            pos += start_part * b.partlen;
            n_file -= (uint32_t)start_part;
            start_part = 0;
        }
        if ((int64_t)n_file > n_to_read - n_read)
            n_file = (uint32_t)(n_to_read - n_read);
        FILE *fd = fopen(f.name.c_str(), "rb");
        if (!fd)
            continue;
        uint32_t got = 0;
        if (fseek(fd, pos, SEEK_SET) == 0)
            got = (uint32_t)fread((char *)dst + n_read * b.partlen, b.partlen, n_file, fd);
        fclose(fd);
        n_read += got;
        if (n_read == n_to_read)
            break;
    }
    return n_read;
}

}  // namespace genpk_host
