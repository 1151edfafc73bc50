// Small device helpers shared by the scan and the gather kernels.
#pragma once

#include <cstdint>

#include "device_types.hpp"

namespace tls {

__device__ __forceinline__ uint32_t divWindow(const DevConfig &c, uint32_t rel) {
    const uint32_t t = __umulhi(c.winMagic, rel);
    return (t + ((rel - t) >> c.winSh1)) >> c.winSh2;
}

// "Terminal" status of a match that starts at item-relative position rel in window winIdx of an
// item of n bases (src/teloscope.cpp:403-406 and :451-459; note the <= in the per-match test).
__device__ __forceinline__ bool isTerminalMatch(const DevConfig &c, uint32_t rel, uint32_t winIdx, uint32_t n) {
    const uint32_t T = c.terminalLimit;
    const uint64_t wStart = static_cast<uint64_t>(winIdx) * c.windowSize;
    uint64_t wEnd = wStart + c.windowSize;
    if (wEnd > n) wEnd = n;
    const uint32_t termEnd = n > T ? n - T : 0;
    if (wEnd <= T || wStart >= termEnd) return true;
    if (wStart > T && wEnd < termEnd) return false;
    return rel <= T || rel >= termEnd;
}

// Window-row kernel (scan_rows.cu): the part [lo, hi) of a contig of n bases that its SPARSE work units cover.  A unit is
// kRowsPerUnit consecutive windows; it is sparse when none of its windows touches a terminal zone (the first / last
// terminalLimit bases, src/teloscope.cpp:403-406) and all of them are full: rel > T and rel + unit span <= n - T.  Only the
// canonical matches of a sparse unit become match records; everywhere else (dense units, other kernels) every match does.
// lo == hi == 0xFFFFFFFF: the contig has no sparse unit.
__host__ __device__ inline void sparseRange(uint32_t n, uint32_t T, uint32_t W, uint32_t &lo, uint32_t &hi) {
    const unsigned long long span = static_cast<unsigned long long>(kRowsPerUnit) * W;
    const unsigned long long termEnd = n > T ? n - T : 0;
    const unsigned long long l = (T / span + 1) * span, h = (termEnd / span) * span;
    if (l < h) { lo = static_cast<uint32_t>(l); hi = static_cast<uint32_t>(h); }
    else lo = hi = 0xFFFFFFFFu;
}

// One overflow chunk of `chunk` records from the bump-allocated arena.  The cursor is 64 bits wide, so the host always
// reads the true demand (a dense batch cannot wrap it); the returned base saturates below kNoOverflow and every write
// is bounds-checked against the arena capacity (<= kMaxOvRecords) in 64 bits.
constexpr uint32_t kMaxOvRecords = 0xF0000000u;
__device__ __forceinline__ uint32_t takeOverflowChunk(unsigned int *cursor, uint32_t chunk) {
    const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long *>(cursor), static_cast<unsigned long long>(chunk));
    return old < kMaxOvRecords ? static_cast<uint32_t>(old) : kMaxOvRecords;
}

// tls_window8 (include/teloscope_b200.h): 10-bit base counts, 8-bit match counts; a window with a match count >= 255 also
// leaves its exact counts in the exception list.
__device__ __forceinline__ uint2 packWindow8(uint32_t window, uint32_t cA, uint32_t cC, uint32_t cG, uint32_t cT, uint32_t can, uint32_t all,
                                             uint32_t fwd, uint4 *winExc, unsigned int *winExcCount) {
    if (max(max(can, all), fwd) >= 255u) winExc[atomicAdd(winExcCount, 1u)] = make_uint4(window, can, all, fwd);
    return make_uint2(cA | (cC << 10) | (cG << 20) | (cT << 30), (cT >> 2) | (min(can, 255u) << 8) | (min(all, 255u) << 16) | (min(fwd, 255u) << 24));
}

}  // namespace tls
