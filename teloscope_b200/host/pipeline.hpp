// Shared pieces of the teloscope-b200 host pipeline (fasta_pipeline.cpp: flags, FASTA / FASTQ front ends, writers;
// bam_subset.cpp: BGZF + BAM; gfa_pipeline.cpp: GFA annotation): options, error exits, input mapping, thread helpers.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

#include "teloscope_b200.h"

namespace tlsb {

struct Options {
    std::string inSequence, inSequenceName, inSequencePrefix, outRoute;
    bool outRouteSet = false;
    std::string canonicalFwd = "CCCTAA", canonicalRev = "TTAGGG";
    std::vector<std::string> rawPatterns;
    bool hasInputPatterns = false;
    uint32_t windowSize = 1000, step = 1000, terminalLimit = 50000;
    uint8_t editDistance = 1;
    unsigned short maxMatchDist = 50, minBlockLen = 300, maxBlockDist = 500, minBlockCounts = 2;
    float minBlockDensity = 0.5f;
    bool outWinRepeats = false, outGC = false, outEntropy = false, outMatches = false, outITS = false;
    bool ultraFastMode = true, manualCuration = false;
    bool fastqSubset = false, bamSubset = false, outFasta = false, minBlockLenSet = false;
    int cmdFlag = 0;
    unsigned threads = 0;                       // -j: host threads of the parse / formatting stages (0 = all cores)
    int gpus = 1;                               // --gpus: devices one job is spread over (one context and feeder thread each)
};

// Leaves without running static destructors: the CUDA context may still be coming up on its own thread, and tearing
// the runtime down underneath it crashes.  Everything written so far is flushed first.
[[noreturn]] inline void quit(int status) {
    fflush(stdout);
    fflush(stderr);
    _exit(status);
}
[[noreturn]] inline void die(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
inline void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    quit(EXIT_FAILURE);
}
[[noreturn]] inline void unsupported(const char *what) {
    fprintf(stderr, "teloscope-b200: %s is not part of the B200 pipeline (use the reference binary for it).\n", what);
    quit(2);
}
inline void check(int rc, const char *what) {
    if (rc != TLS_OK) die("Error: %s failed: %s\n", what, tls_last_error());
}

struct StageTimer {                       // TLS_PIPE_TRACE=1: wall time of the pipeline stages on stderr
    bool on = getenv("TLS_PIPE_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    void mark(const char *what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pipe] %-28s %8.1f ms  (total %8.1f ms)\n", what, std::chrono::duration<double, std::milli>(now - last).count(),
                std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

// Input bytes: a plain file is mapped (no copy); gzip files and stdin are read into memory.
struct InputData {
    const char *p = nullptr;
    size_t n = 0;
    std::string owned;
    void *map = nullptr;
    ~InputData() { if (map) munmap(map, n); }
};
inline bool mapFile(const std::string &path, InputData &in) {      // false: cannot open
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st {};
    fstat(fd, &st);
    if (st.st_size > 0) {
        void *m = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, static_cast<size_t>(st.st_size), MADV_SEQUENTIAL);
            in.map = m; in.p = static_cast<const char *>(m); in.n = static_cast<size_t>(st.st_size);
            close(fd);
            return true;
        }
        std::vector<char> buf(1 << 22);
        for (ssize_t got; (got = read(fd, buf.data(), buf.size())) > 0;) in.owned.append(buf.data(), static_cast<size_t>(got));
        in.p = in.owned.data(); in.n = in.owned.size();
    }
    close(fd);
    return true;
}
// a plain file is mapped (no copy), a gzip file is inflated by zlib into memory
inline void loadInput(const std::string &path, InputData &in) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) die("Error: Cannot open input file '%s'.\n", path.c_str());
    unsigned char magic[2] = {0, 0};
    const ssize_t got = pread(fd, magic, 2, 0);
    struct stat st {};
    fstat(fd, &st);
    if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) && st.st_size > 0) {
        void *m = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, static_cast<size_t>(st.st_size), MADV_SEQUENTIAL);
            in.map = m; in.p = static_cast<const char *>(m); in.n = static_cast<size_t>(st.st_size);
            close(fd);
            return;
        }
    }
    close(fd);
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) die("Error: Cannot open input file '%s'.\n", path.c_str());
    gzbuffer(f, 1 << 20);
    std::vector<char> buf(1 << 24);
    for (;;) {
        const int n = gzread(f, buf.data(), static_cast<unsigned>(buf.size()));
        if (n < 0) die("Error: Cannot read input file '%s'.\n", path.c_str());
        if (n == 0) break;
        in.owned.append(buf.data(), static_cast<size_t>(n));
    }
    gzclose(f);
    in.p = in.owned.data(); in.n = in.owned.size();
}

inline void readAll(FILE *f, InputData &in) {
    std::vector<char> buf(1 << 22);
    for (size_t got; (got = fread(buf.data(), 1, buf.size(), f)) > 0;) in.owned.append(buf.data(), got);
    in.p = in.owned.data(); in.n = in.owned.size();
}

// `threads` copies of f side by side (f pulls its work from an atomic counter)
template <class F>
inline void runThreads(size_t threads, F f) {
    if (threads <= 1) { f(); return; }
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; ++t) pool.emplace_back(f);
    for (auto &th : pool) th.join();
}
// dynamic parallel loop over [0, n): the host side of the pipeline (parse, formatting) is per-record / per-chunk work
template <class F>
inline void parallelFor(size_t n, unsigned threads, F f) {
    if (n == 0) return;
    threads = static_cast<unsigned>(std::min<size_t>(std::max(1u, threads), n));
    if (threads == 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    runThreads(threads, [&]() { for (size_t i; (i = next.fetch_add(1)) < n;) f(i); });
}

int runBamSubset(const Options &o, const tls_config &baseCfg);      // bam_subset.cpp
bool isGfaPath(const std::string &path);                            // gfa_pipeline.cpp
int runGfa(const Options &o, const tls_config &baseCfg);
int normalizeGfa(const std::string &inPath, const std::string &outPath);   // the graph read and written back without annotation (no device)

}  // namespace tlsb
