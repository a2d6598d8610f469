// Positional reads of many byte ranges at once: the pieces of a block that live in different files of a snapshot
// (and the slices of a large piece) are fetched by a few threads with pread(), each straight to its place in the
// destination buffer.  The reference reads them one after the other through one FILE* (gadgetreader.cpp:471-557,
// bigfile.c:714-801); what lands in the buffer is the same.
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <thread>

#include "snapshot.hpp"

namespace genpk_host {

static bool pread_all(int fd, char *dst, int64_t bytes, int64_t offset)
{
    while (bytes > 0) {
        const ssize_t got = pread(fd, dst, (size_t)bytes, (off_t)offset);
        if (got <= 0)
            return false;
        dst += got;
        offset += got;
        bytes -= got;
    }
    return true;
}

bool read_segments(const std::vector<ReadSeg> &segs, int max_threads, std::string *error)
{
    // slices of at most 32 MB, so that one big file is read by several threads too
    constexpr int64_t SLICE = (int64_t)32 << 20;
    struct Piece {
        size_t seg;
        int64_t offset, bytes;
    };
    std::vector<Piece> pieces;
    for (size_t s = 0; s < segs.size(); s++)
        for (int64_t done = 0; done < segs[s].bytes; done += SLICE)
            pieces.push_back({s, done, std::min(SLICE, segs[s].bytes - done)});
    if (pieces.empty())
        return true;
    std::vector<int> fds(segs.size(), -1);
    for (size_t s = 0; s < segs.size(); s++) {
        if (segs[s].bytes <= 0)
            continue;
        fds[s] = open(segs[s].path.c_str(), O_RDONLY);
        if (fds[s] < 0) {
            if (error) *error = "cannot open " + segs[s].path;
            for (int fd : fds)
                if (fd >= 0) close(fd);
            return false;
        }
    }
    std::atomic<size_t> next(0);
    std::atomic<int> failed(-1);
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= pieces.size() || failed.load() >= 0)
                return;
            const Piece &p = pieces[i];
            const ReadSeg &g = segs[p.seg];
            if (!pread_all(fds[p.seg], (char *)g.dst + p.offset, p.bytes, g.offset + p.offset))
                failed.store((int)p.seg);
        }
    };
    int nthreads = (int)std::min<size_t>(pieces.size(), (size_t)std::max(1, max_threads));
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++)
        pool.emplace_back(work);
    work();
    for (auto &t : pool)
        t.join();
    for (int fd : fds)
        if (fd >= 0) close(fd);
    if (failed.load() >= 0) {
        if (error) *error = "short read in " + segs[(size_t)failed.load()].path;
        return false;
    }
    return true;
}

int read_threads()
{
    static int n = 0;
    if (n == 0) {
        const char *env = getenv("GENPK_READ_THREADS");
        n = env ? atoi(env) : 0;
        if (n <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            n = hw ? (int)std::min(hw, 8u) : 4;
        }
    }
    return n;
}

}  // namespace genpk_host
