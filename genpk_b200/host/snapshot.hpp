// Snapshot readers of the gen-pk host (Gadget-I/II binary and MP-Gadget bigfile).
// They hand the deposit float32 xyz triples (and float32 masses) exactly as the
// reference's adapters do (read_fieldize.cpp:18-97, read_fieldize_bigfile.cpp:64-125).
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace genpk_host {

constexpr int N_TYPE = 6;

// One byte range of one file and where it goes (parallel_read.cpp).
struct ReadSeg {
    std::string path;
    int64_t offset = 0, bytes = 0;
    void *dst = nullptr;
};
// All of them, fetched by up to max_threads threads with pread(); false (and *error) on a missing file or a short read.
bool read_segments(const std::vector<ReadSeg> &segs, int max_threads, std::string *error);
int read_threads();                                      // GENPK_READ_THREADS, default min(hardware threads, 8)

// The 256-byte Gadget header (layout of GadgetReader/gadgetheader.h:18-76).
#pragma pack(push, 1)
struct GadgetHeader {
    uint32_t npart[N_TYPE];
    double mass[N_TYPE];
    double time, redshift;
    int32_t flag_sfr, flag_feedback;
    uint32_t npartTotal[N_TYPE];
    int32_t flag_cooling, num_files;
    double BoxSize, Omega0, OmegaLambda, HubbleParam;
    int32_t flag_stellarage, flag_metals;
    uint32_t NallHW[N_TYPE];
    char fill[256 - 6 * 4 - 6 * 8 - 2 * 8 - 2 * 4 - 6 * 4 - 2 * 4 - 4 * 8 - 2 * 4 - 6 * 4];
};
#pragma pack(pop)
static_assert(sizeof(GadgetHeader) == 256, "Gadget header is 256 bytes");

struct GadgetBlock {
    int64_t start = 0;      // file offset of the first data byte
    int64_t length = 0;     // bytes of data
    int partlen = 0;        // bytes per particle
};

struct GadgetFile {
    std::string name;
    GadgetHeader header = {};
    std::map<std::string, GadgetBlock> blocks;
};

// A (multi-file) Gadget format II snapshot: base, base.0 ... base.(num_files-1).
class GadgetSnapshot {
public:
    explicit GadgetSnapshot(const std::string &base);
    bool ok() const { return !files_.empty(); }
    int num_files() const { return (int)files_.size(); }
    const GadgetHeader &header() const { return files_[0].header; }
    int64_t npart(int type) const;                       // from the header of file 0, as GSnap::GetNpart
    bool has_block(const std::string &name) const;
    int64_t block_bytes(const std::string &name) const;  // summed over files
    int64_t block_parts(const std::string &name) const;
    // Same contract and the same offset arithmetic as GSnap::GetBlock
    // (gadgetreader.cpp:471-557), including what it does when `skip_type` names
    // types the block does not hold: callers get the bytes the
    // reference's reader would have handed to fieldize().
    int64_t get_block(const std::string &name, void *dst, int64_t n_to_read, int64_t start_part, int skip_type) const;
    int64_t get_block_sequential(const std::string &name, void *dst, int64_t n_to_read, int64_t start_part, int skip_type) const;
    const std::string &error() const { return error_; }

private:
    bool scan_file(const std::string &path, GadgetFile *out);
    std::vector<GadgetFile> files_;
    std::string error_;
};

// ---- bigfile (directory-per-block column store, bigfile/src/bigfile.c) ----------------
struct BigBlockInfo {
    std::string dir;
    std::string dtype;             // e.g. "<f8"
    int nmemb = 0;
    std::vector<int64_t> file_rows;
    int64_t rows = 0;
    int itemsize = 0;
};

class BigfileSnapshot {
public:
    explicit BigfileSnapshot(const std::string &dir);
    bool ok() const { return ok_; }                      // directory with a Header block (is_bigfile)
    // attributes of Header/attr-v2, converted to double / int64
    bool attr_f8(const std::string &name, double *out, int n) const;
    bool attr_i8(const std::string &name, int64_t *out, int n) const;
    bool open_block(const std::string &name, BigBlockInfo *info) const;
    // rows [first, first+count) of a block converted to float32 (through double when the
    // stored type is f8, as read_fieldize_bigfile.cpp:82-95 does)
    bool read_f32(const BigBlockInfo &b, int64_t first, int64_t count, float *dst) const;
    // rows [first, first+count) of a little-endian f8 block as stored (the GPU narrows them:
    // genpk_deposit_f64); false for any other dtype
    bool read_f64_raw(const BigBlockInfo &b, int64_t first, int64_t count, double *dst) const;
    const std::string &error() const { return error_; }

private:
    struct Attr {
        std::string dtype;
        int nmemb = 0;
        std::vector<unsigned char> raw;
    };
    std::string dir_;
    bool ok_ = false;
    std::map<std::string, Attr> attrs_;
    mutable std::string error_;
};

}  // namespace genpk_host
