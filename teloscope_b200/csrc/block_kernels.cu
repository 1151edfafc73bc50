// Kernels that turn the ordered match list into what SegmentData holds besides the matches:
//   itemOffsetsKernel   exclusive scan of per-item match counts
//   terminalKernel      Teloscope::getTerminalBlocks (src/teloscope.cpp:29-176), one walk per
//                       (item, strand), Phase 2 merged on the fly
//   itsKernel           Teloscope::getInterstitialBlocks (src/teloscope.cpp:179-256): every run
//                       start walks its run
//   selectKernels       ordered selection of canonical / terminal non-canonical matches (-m)
#include <cuda_runtime.h>

#include <cstdint>

#include "device_common.cuh"
#include "kernels.hpp"

namespace tls {

// ------------------------------------------------------------------------------------------------
// Exclusive scan of itemCounts[i].all -> itemBegin[0..n]   (single block, n up to millions)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) itemOffsetsKernel(const DevItemCounts *counts, uint32_t n,
                                                          unsigned long long *itemBegin) {
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + tid;
        const unsigned long long v = i < n ? counts[i].all : 0;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warpSums[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += u;
            }
            warpSums[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long excl = carry + warpSums[warp] + incl - v;
        if (i < n) itemBegin[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) itemBegin[n] = carry;
}

// ------------------------------------------------------------------------------------------------
// Slot arena -> one ordered match list.
//   tileScanKernel: per-tile totals (sum of the 32 row slots) and their exclusive prefix
//   gatherKernel:   one block per tile copies its slots (and overflow chunks) to the list
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tileScanKernel(const GatherParams g) {
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < g.nTiles; base += 1024) {
        const uint32_t i = base + tid;
        unsigned long long v = 0;
        if (i < g.nTiles) {
            const uint4 *c = reinterpret_cast<const uint4 *>(g.slotCounts + static_cast<size_t>(i) * kTileRows);
#pragma unroll
            for (int k = 0; k < kTileRows / 4; ++k) { const uint4 x = c[k]; v += x.x; v += x.y; v += x.z; v += x.w; }
        }
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warpSums[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += u;
            }
            warpSums[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long excl = carry + warpSums[warp] + incl - v;
        if (i < g.nTiles) g.tileBase[i] = excl;
        __syncthreads();
        if (tid == 1023) carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) g.tileBase[g.nTiles] = carry;
}

// One block per tile.  Thread k walks the k-th contiguous share of the tile's records in genome
// order: copies them to the ordered list (adding the terminal flag) and accumulates, while the
// key stays the same, per-item counters and per-window covered bp; shared atomics on key change,
// one flush to global memory at the end.
constexpr int kGatherThreads = 128;
constexpr int kGatherSlots = 128;

__global__ void __launch_bounds__(kGatherThreads) gatherKernel(const GatherParams g) {
    const uint32_t tile = blockIdx.x;
    const int tid = threadIdx.x;
    __shared__ uint32_t sOff[kTileRows + 1];
    __shared__ DevRegion sRegion[kRegionCache];
    __shared__ uint32_t sRegCnt[kRegionCache][5];
    __shared__ uint32_t sWin[kGatherSlots][4];
    __shared__ uint32_t sFirstWin, sNReg;
    const DevConfig &c = g.cfg;
    const DevTile dt = g.tiles[tile];
    if (tid == 0) {
        uint32_t acc = 0;
        for (int w = 0; w < kTileRows; ++w) { sOff[w] = acc; acc += g.slotCounts[static_cast<size_t>(tile) * kTileRows + w]; }
        sOff[kTileRows] = acc;
    }
    if (tid < 32) {
        const uint64_t tileEnd = dt.bufStart + dt.len;
        const uint32_t idx = min(dt.firstRegion + static_cast<uint32_t>(tid), g.nRegions - 1);
        const DevRegion r = g.regions[idx];
        const bool in = dt.firstRegion + tid < g.nRegions && r.bufStart < tileEnd;
        const uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, in));
        if (in) sRegion[tid] = r;
        if (tid == 0) {
            sNReg = n;
            const uint64_t s = dt.bufStart > r.bufStart ? dt.bufStart : r.bufStart;
            sFirstWin = (c.mode == 0 && n > 0) ? r.winBase + divWindow(c, r.relStart + static_cast<uint32_t>(s - r.bufStart)) : 0;
        }
    }
    for (int i = tid; i < kRegionCache * 5; i += kGatherThreads) (&sRegCnt[0][0])[i] = 0;
    for (int i = tid; i < kGatherSlots * 4; i += kGatherThreads) (&sWin[0][0])[i] = 0;
    __syncthreads();
    const uint32_t total = sOff[kTileRows];
    if (total == 0) return;
    const unsigned long long base = g.tileBase[tile];
    const uint32_t firstWin = sFirstWin;
    const uint32_t share = (total + kGatherThreads - 1) / kGatherThreads;
    uint32_t i = tid * share;
    const uint32_t end = min(total, i + share);

    // running accumulators
    uint32_t curReg = 0xFFFFFFFFu, nAll = 0, nFwd = 0, nCan = 0, nNct = 0;
    uint32_t curWin = 0xFFFFFFFFu, canCov = 0, nonCov = 0, fwdCov = 0, revCov = 0;
    auto flushReg = [&]() {
        if (curReg != 0xFFFFFFFFu && nAll) {
            uint32_t *r = sRegCnt[curReg];
            atomicAdd(r + 0, nAll); atomicAdd(r + 1, nFwd); atomicAdd(r + 2, nAll - nFwd);
            if (nCan) atomicAdd(r + 3, nCan);
            if (nNct) atomicAdd(r + 4, nNct);
        }
        nAll = nFwd = nCan = nNct = 0;
    };
    auto flushWin = [&]() {
        if (curWin != 0xFFFFFFFFu) {
            const uint32_t s = curWin - firstWin;
            uint32_t *w = s < kGatherSlots ? sWin[s] : g.windows + static_cast<size_t>(curWin) * 8 + 4;
            if (canCov) atomicAdd(w + 0, canCov);
            if (nonCov) atomicAdd(w + 1, nonCov);
            if (fwdCov) atomicAdd(w + 2, fwdCov);
            if (revCov) atomicAdd(w + 3, revCov);
        }
        canCov = nonCov = fwdCov = revCov = 0;
    };

    if (i < end) {
        int w = 0;
        while (sOff[w + 1] <= i) ++w;
        for (; i < end; ++i) {
            while (sOff[w + 1] <= i) ++w;
            const size_t slotIdx = static_cast<size_t>(tile) * kTileRows + w;
            const uint32_t k = i - sOff[w];
            DevMatch m;
            if (k < g.slotCap) m = g.slots[slotIdx * g.slotCap + k];
            else {
                const unsigned long long o = static_cast<unsigned long long>(g.slotOv[slotIdx]) + (k - g.slotCap);
                if (o >= g.ovCap) continue;                     // overflow arena exhausted: the host reruns
                m = g.ovArena[o];
            }
            if (m.region != kRegionAccounted) {
                const DevRegion &r = sRegion[m.region];
                const bool fwd = m.flags & kMatchFwd, can = m.flags & kMatchCan;
                bool term = false;
                if (c.mode == 0) {
                    const uint32_t winIdx = divWindow(c, m.pos);
                    term = isTerminalMatch(c, m.pos, winIdx, r.itemLen);
                    const uint32_t win = r.winBase + winIdx;
                    if (win != curWin) { flushWin(); curWin = win; }
                    if (can) canCov += m.len; else nonCov += m.len;
                    if (fwd) fwdCov += m.len; else revCov += m.len;
                }
                if (m.region != curReg) { flushReg(); curReg = m.region; }
                ++nAll; nFwd += fwd; nCan += can; nNct += (!can && term);
                if (term) m.flags |= kMatchTerm;
            }
            m.region = 0;
            const unsigned long long o = base + i;
            if (o < g.matchCap) g.matches[o] = m;
        }
    }
    flushReg();
    flushWin();
    __syncthreads();
    for (int k = tid; k < kGatherSlots * 4; k += kGatherThreads) {
        const uint32_t v = (&sWin[0][0])[k];
        if (v) atomicAdd(&g.windows[static_cast<size_t>(firstWin + (k >> 2)) * 8 + 4 + (k & 3)], v);
    }
    for (int k = tid; k < static_cast<int>(sNReg) * 5; k += kGatherThreads) {
        const uint32_t v = (&sRegCnt[0][0])[k];
        if (v) atomicAdd(reinterpret_cast<uint32_t *>(g.itemCounts + sRegion[k / 5].item) + (k % 5), v);
    }
}

cudaError_t launchTileScan(const GatherParams &g, cudaStream_t s) {
    tileScanKernel<<<1, 1024, 0, s>>>(g);
    return cudaGetLastError();
}
cudaError_t launchGather(const GatherParams &g, cudaStream_t s) {
    if (g.nTiles == 0) return cudaSuccess;
    gatherKernel<<<g.nTiles, kGatherThreads, 0, s>>>(g);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Terminal blocks
// ------------------------------------------------------------------------------------------------
namespace {

struct Run {
    uint32_t start, end, prev, n, nFwd, nCan, cov, fwdCov, canCov;
    __device__ void begin(const DevMatch &m) {
        const uint32_t f = m.flags & kMatchFwd ? 1u : 0u, c = m.flags & kMatchCan ? 1u : 0u;
        start = m.pos; end = m.pos + m.len; prev = m.pos; n = 1; nFwd = f; nCan = c;
        cov = m.len; fwdCov = f * m.len; canCov = c * m.len;
    }
    __device__ void absorb(const DevMatch &m) {
        const uint32_t f = m.flags & kMatchFwd ? 1u : 0u, c = m.flags & kMatchCan ? 1u : 0u;
        ++n; nFwd += f; nCan += c; cov += m.len; fwdCov += f * m.len; canCov += c * m.len; prev = m.pos;
    }
};

__device__ void emitBlock(const BlockParams &p, const DevBlock &b) {
    const unsigned int slot = atomicAdd(p.blockCount, 1u);
    if (slot < p.blockCap) p.blocks[slot] = b;
}

}  // namespace

// One thread per (item, arm).  arm 0 = p (forward-strand matches from the start),
// arm 1 = q (reverse-strand matches from the end).
__global__ void terminalKernel(const BlockParams p) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= 2u * p.nItems) return;
    const uint32_t item = gid >> 1;
    const bool fromStart = (gid & 1u) == 0;
    const DevConfig &c = p.cfg;
    const uint32_t n = p.itemLen[item];
    const bool readMode = c.mode == 2;
    const uint32_t T = readMode ? 0x7FFFFFFFu : c.terminalLimit;      // UINT32_MAX / 2, src/read-filter.cpp:19
    uint32_t boundary = fromStart ? 0u : n;
    const DevItemCounts cnt = p.itemCounts[item];
    const uint32_t strandCount = fromStart ? cnt.fwd : cnt.rev;
    uint32_t emitted = 0;

    if (strandCount >= 2) {                                           // src/teloscope.cpp:646,649
        const long long begin = static_cast<long long>(p.itemBegin[item]);
        const long long end = static_cast<long long>(p.itemBegin[item + 1]);
        const uint8_t want = fromStart ? kMatchFwd : 0;
        auto inZone = [&](uint32_t rel) {                             // :48-52
            if (n <= T) return true;
            return fromStart ? rel < T : rel >= n - T;
        };
        // Phase 2 state: the block being extended
        bool haveCur = false;
        DevBlock cur{};
        auto finishCur = [&]() {                                      // :132-145
            if (cur.len >= c.minBlockLen) {
                cur.item = item;
                cur.order = (fromStart ? 0u : 0x80000000u) | emitted;
                cur.label = fromStart ? 'p' : 'q';
                const uint64_t relEnd = static_cast<uint64_t>(cur.start) + cur.len;
                const uint64_t left = cur.start, right = relEnd <= n ? n - relEnd : 0;
                cur.validOr = fromStart ? left <= right : left >= right;
                emitBlock(p, cur);
                ++emitted;
                boundary = fromStart ? cur.start + cur.len : cur.start;
            }
        };
        auto feedSub = [&](const Run &r) {                            // Phase 1 filter :75-92, Phase 2 :147-172
            const float rhs = __fmul_rn(c.minBlockDensity, __ull2float_rn(static_cast<unsigned long long>(r.end - r.start)));
            if (!(r.n >= c.minBlockCounts && r.nCan > 0 && __uint2float_rn(r.canCov) >= rhs)) return;
            DevBlock nx{};
            nx.start = r.start; nx.len = r.end - r.start; nx.counts = r.n; nx.fwdCount = r.nFwd;
            nx.canCount = r.nCan; nx.totalCov = r.cov; nx.fwdCov = r.fwdCov; nx.canCov = r.canCov;
            if (!haveCur) { cur = nx; haveCur = true; return; }
            const uint64_t gap = fromStart ? static_cast<uint64_t>(nx.start) - (static_cast<uint64_t>(cur.start) + cur.len)
                                           : static_cast<uint64_t>(cur.start) - (static_cast<uint64_t>(nx.start) + nx.len);
            if (gap <= c.maxBlockDist) {
                if (fromStart) cur.len = nx.start + nx.len - cur.start;
                else { cur.len = cur.start + cur.len - nx.start; cur.start = nx.start; }
                cur.counts += nx.counts; cur.fwdCount += nx.fwdCount; cur.canCount += nx.canCount;
                cur.totalCov += nx.totalCov; cur.fwdCov += nx.fwdCov; cur.canCov += nx.canCov;
            } else {
                finishCur();
                cur = nx;
            }
        };

        Run run{};
        bool inRun = false;
        const long long step = fromStart ? 1 : -1;
        for (long long i = fromStart ? begin : end - 1; i != (fromStart ? end : begin - 1); i += step) {
            const DevMatch m = p.matches[i];
            if ((m.flags & kMatchFwd) != want) continue;
            if (!inRun) {
                if (!inZone(m.pos)) break;
                run.begin(m);
                inRun = true;
                continue;
            }
            const uint32_t gap = fromStart ? m.pos - run.prev : run.prev - m.pos;
            if (gap <= c.maxMatchDist) {
                if (fromStart) run.end = m.pos + m.len; else run.start = m.pos;
                run.absorb(m);
            } else {
                feedSub(run);
                inRun = false;
                if (!inZone(m.pos)) break;
                run.begin(m);
                inRun = true;
            }
        }
        if (inRun) feedSub(run);
        if (haveCur) finishCur();
    }
    if (fromStart) p.fwdBoundary[item] = boundary; else p.revBoundary[item] = boundary;
    if (readMode && emitted) p.pass[item] = 1;
}

// ------------------------------------------------------------------------------------------------
// Interstitial blocks.  One thread per 64 consecutive matches; a thread that finds a run start
// among its matches walks that run to its end.
// ------------------------------------------------------------------------------------------------
__global__ void itsKernel(const BlockParams p, unsigned long long nMatches) {
    const unsigned long long first = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 64ull;
    if (first >= nMatches) return;
    const DevConfig &c = p.cfg;
    // item of the first match: last i with itemBegin[i] <= first
    uint32_t lo = 0, hi = p.nItems;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.itemBegin[mid] <= first) lo = mid + 1; else hi = mid;
    }
    uint32_t item = lo - 1;
    unsigned long long itemEnd = p.itemBegin[item + 1];
    const unsigned long long last = first + 64 < nMatches ? first + 64 : nMatches;
    for (unsigned long long i = first; i < last; ++i) {
        while (i >= itemEnd) { ++item; itemEnd = p.itemBegin[item + 1]; }
        const unsigned long long itemBegin = p.itemBegin[item];
        const uint32_t fwdB = p.fwdBoundary[item], revB = p.revBoundary[item];
        // src/teloscope.cpp:653: needs fwdBoundary < revBoundary and >= 2 matches in the segment
        if (!(fwdB < revB) || itemEnd - itemBegin < 2) { i = itemEnd - 1; continue; }
        const DevMatch m = p.matches[i];
        if (m.pos < fwdB || m.pos >= revB) continue;
        bool isStart = true;                                  // first match at/after fwdBoundary starts a run
        if (i > itemBegin) {
            const DevMatch prev = p.matches[i - 1];
            if (prev.pos >= fwdB) isStart = m.pos - prev.pos > c.maxMatchDist;
        }
        if (!isStart) continue;
        Run r{};
        r.begin(m);
        for (unsigned long long j = i + 1; j < itemEnd; ++j) {
            const DevMatch x = p.matches[j];
            if (x.pos >= revB || x.pos - r.prev > c.maxMatchDist) break;
            r.end = x.pos + x.len;
            r.absorb(x);
        }
        const uint32_t len = r.end - r.start;
        // computeBlockLabel, include/teloscope.h:217-222 (float32)
        const float ratio = __fdiv_rn(__fmul_rn(__uint2float_rn(r.nFwd), 100.0f), __uint2float_rn(r.n));
        const char label = ratio > 66.6f ? 'p' : (ratio < 33.3f ? 'q' : 'b');
        if (len >= c.itsMinLen && r.nCan >= 4 && !(label == 'b' && r.nFwd < 2 && (r.n - r.nFwd) < 2)) {
            DevBlock b{};
            b.item = item; b.order = 0; b.start = r.start; b.len = len; b.counts = r.n; b.fwdCount = r.nFwd;
            b.canCount = r.nCan; b.totalCov = r.cov; b.fwdCov = r.fwdCov; b.canCov = r.canCov;
            b.validOr = 1; b.label = label; b.kind = 1;
            emitBlock(p, b);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Ordered selection for -m: canonical matches, and non-canonical matches flagged terminal.
// Pass 1 counts per 2048-match chunk, a single block scans the chunk counts, pass 2 writes.
// ------------------------------------------------------------------------------------------------
constexpr int kSelChunk = 2048;

__global__ void __launch_bounds__(256) selectCountKernel(const DevMatch *m, unsigned long long n, uint32_t *chunkCounts) {
    const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kSelChunk;
    uint32_t nc = 0, nn = 0;
    for (int i = threadIdx.x; i < kSelChunk; i += 256) {
        const unsigned long long k = base + i;
        if (k < n) {
            const uint8_t f = m[k].flags;
            nc += (f & kMatchCan) != 0;
            nn += (f & (kMatchCan | kMatchTerm)) == kMatchTerm;
        }
    }
    nc = __reduce_add_sync(0xFFFFFFFFu, nc);
    nn = __reduce_add_sync(0xFFFFFFFFu, nn);
    __shared__ uint32_t s[2];
    if (threadIdx.x < 2) s[threadIdx.x] = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s[0], nc); atomicAdd(&s[1], nn); }
    __syncthreads();
    if (threadIdx.x < 2) chunkCounts[2 * blockIdx.x + threadIdx.x] = s[threadIdx.x];
}

// in place: chunkCounts[2*i+k] -> exclusive prefix; totals[k] = sum
__global__ void __launch_bounds__(1024) selectScanKernel(uint32_t *chunkCounts, uint32_t nChunks, unsigned long long *totals) {
    __shared__ unsigned long long warpSums[2][32];
    __shared__ unsigned long long carry[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) carry[tid] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nChunks; base += 1024) {
        const uint32_t i = base + tid;
        unsigned long long v[2], incl[2];
        for (int k = 0; k < 2; ++k) {
            v[k] = i < nChunks ? chunkCounts[2 * i + k] : 0;
            incl[k] = v[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, incl[k], d);
                if (lane >= d) incl[k] += u;
            }
            if (lane == 31) warpSums[k][warp] = incl[k];
        }
        __syncthreads();
        if (warp < 2) {
            unsigned long long w = warpSums[warp][lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= d) wi += u;
            }
            warpSums[warp][lane] = wi - w;
        }
        __syncthreads();
        unsigned long long excl[2];
        for (int k = 0; k < 2; ++k) {
            excl[k] = carry[k] + warpSums[k][warp] + incl[k] - v[k];
            // prefix fits 32 bits only if the list does; keep 32-bit chunk offsets relative to 0
            if (i < nChunks) chunkCounts[2 * i + k] = static_cast<uint32_t>(excl[k]);
        }
        __syncthreads();
        if (tid == 1023) { carry[0] = excl[0] + v[0]; carry[1] = excl[1] + v[1]; }
        __syncthreads();
    }
    if (tid < 2) totals[tid] = carry[tid];
}

// Final host-visible match record (same layout as tls_match in include/teloscope_b200.h).
struct OutMatch {
    unsigned long long pos;
    uint16_t len;
    uint8_t isForward, isCanonical;
    uint32_t pad;
};
static_assert(sizeof(OutMatch) == 16, "tls_match layout");

__global__ void __launch_bounds__(32) selectWriteKernel(const DevMatch *m, unsigned long long n, const uint32_t *chunkOffsets,
                                                        const unsigned long long *itemBegin, const unsigned long long *absPos,
                                                        uint32_t nItems, OutMatch *outCan, uint32_t capCan, OutMatch *outNon,
                                                        uint32_t capNon) {
    // one warp walks the chunk in order so the output order equals the input order
    const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kSelChunk;
    const int lane = threadIdx.x;
    uint32_t oc = chunkOffsets[2 * blockIdx.x], on = chunkOffsets[2 * blockIdx.x + 1];
    // item of the chunk's first match: last i with itemBegin[i] <= base
    uint32_t lo = 0, hi = nItems;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (itemBegin[mid] <= base) lo = mid + 1; else hi = mid;
    }
    uint32_t item = lo - 1;
    for (int i = 0; i < kSelChunk; i += 32) {
        const unsigned long long k = base + i + lane;
        DevMatch x{};
        bool isC = false, isN = false;
        unsigned long long ap = 0;
        if (k < n) {
            x = m[k];
            isC = (x.flags & kMatchCan) != 0;
            isN = (x.flags & (kMatchCan | kMatchTerm)) == kMatchTerm;
            if (isC || isN) {
                uint32_t it = item;
                while (k >= itemBegin[it + 1]) ++it;
                ap = absPos[it];
            }
        }
        const uint32_t bc = __ballot_sync(0xFFFFFFFFu, isC), bn = __ballot_sync(0xFFFFFFFFu, isN);
        const uint32_t below = (1u << lane) - 1u;
        OutMatch o;
        o.pos = ap + x.pos; o.len = x.len; o.isForward = x.flags & kMatchFwd ? 1 : 0; o.isCanonical = isC ? 1 : 0; o.pad = 0;
        if (isC) { const uint32_t q = oc + __popc(bc & below); if (q < capCan) outCan[q] = o; }
        if (isN) { const uint32_t q = on + __popc(bn & below); if (q < capNon) outNon[q] = o; }
        oc += __popc(bc);
        on += __popc(bn);
        // advance the chunk's item cursor past items that end before the next 32 matches
        const unsigned long long nextK = base + i + 32;
        while (item + 1 < nItems && nextK >= itemBegin[item + 1]) ++item;
    }
}

cudaError_t launchItemOffsets(const DevItemCounts *counts, uint32_t n, unsigned long long *itemBegin, cudaStream_t s) {
    itemOffsetsKernel<<<1, 1024, 0, s>>>(counts, n, itemBegin);
    return cudaGetLastError();
}
cudaError_t launchTerminal(const BlockParams &p, cudaStream_t s) {
    const uint32_t threads = 2u * p.nItems;
    if (threads == 0) return cudaSuccess;
    terminalKernel<<<(threads + 127) / 128, 128, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t launchIts(const BlockParams &p, unsigned long long nMatches, cudaStream_t s) {
    if (nMatches == 0) return cudaSuccess;
    const unsigned long long threads = (nMatches + 63) / 64;
    itsKernel<<<static_cast<unsigned int>((threads + 127) / 128), 128, 0, s>>>(p, nMatches);
    return cudaGetLastError();
}

uint32_t selectChunks(unsigned long long n) { return static_cast<uint32_t>((n + kSelChunk - 1) / kSelChunk); }

cudaError_t launchSelect(const DevMatch *m, unsigned long long n, uint32_t *chunkScratch, unsigned long long *totals,
                         const unsigned long long *itemBegin, const unsigned long long *absPos, uint32_t nItems,
                         void *outCan, uint32_t capCan, void *outNon, uint32_t capNon, cudaStream_t s) {
    const uint32_t chunks = selectChunks(n);
    if (chunks == 0) return cudaSuccess;
    selectCountKernel<<<chunks, 256, 0, s>>>(m, n, chunkScratch);
    selectScanKernel<<<1, 1024, 0, s>>>(chunkScratch, chunks, totals);
    selectWriteKernel<<<chunks, 32, 0, s>>>(m, n, chunkScratch, itemBegin, absPos, nItems, static_cast<OutMatch *>(outCan),
                                            capCan, static_cast<OutMatch *>(outNon), capNon);
    return cudaGetLastError();
}

// Read mode with a device-resident batch: drop one trailing '\r' per read (src/read-filter.cpp:38-40).
__global__ void trimCrKernel(const uint8_t *seq, DevRegion *regions, uint32_t nRegions, uint32_t *itemLen) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nRegions) return;
    DevRegion reg = regions[r];
    if (reg.len && seq[reg.bufStart + reg.len - 1] == '\r') {
        regions[r].len = reg.len - 1;
        regions[r].itemLen = reg.itemLen - 1;
        itemLen[reg.item] = reg.itemLen - 1;
    }
}

cudaError_t launchTrimCr(const uint8_t *seq, DevRegion *regions, uint32_t nRegions, uint32_t *itemLen, cudaStream_t s) {
    if (nRegions == 0) return cudaSuccess;
    trimCrKernel<<<(nRegions + 255) / 256, 256, 0, s>>>(seq, regions, nRegions, itemLen);
    return cudaGetLastError();
}

}  // namespace tls
