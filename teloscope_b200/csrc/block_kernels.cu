// Kernels that turn the per-row match slots into what SegmentData holds besides the window counters:
//   tileScanKernel      exclusive prefixes over tiles of (all, canonical, terminal non-canonical) matches
//   gatherKernel        slot arena -> one ordered match list + index lists of the -m selections
//   itemOffsetsKernel   exclusive scan of per-item match counts
//   terminalKernel      Teloscope::getTerminalBlocks (src/teloscope.cpp:29-176), one warp per
//                       (item, arm), Phase 2 merged on the fly
//   itsSeedKernel /     Teloscope::getInterstitialBlocks (src/teloscope.cpp:179-256): a block needs
//   itsRunKernel        >= 4 canonical matches, so only runs that hold canonical matches are walked
//   selectFinalizeKernel  final 16-byte records of the canonical / terminal non-canonical lists (-m)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "device_common.cuh"
#include "kernels.hpp"

namespace tls {

namespace {

// Exclusive scan of up to 1024 values per call inside one block of 1024 threads; returns the
// exclusive prefix of `v` (carry included) and advances the carry.  warpSums: 32 entries.
__device__ __forceinline__ unsigned long long blockExclusive1024(unsigned long long v, unsigned long long *warpSums,
                                                                 unsigned long long *carry) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const unsigned long long excl = *carry + warpSums[warp] + incl - v;
    __syncthreads();
    if (tid == 1023) *carry = excl + v;
    __syncthreads();
    return excl;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Exclusive scan of itemCounts[i].all -> itemBegin[0..n]   (single block, n up to millions)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) itemOffsetsKernel(const DevItemCounts *counts, uint32_t n,
                                                          unsigned long long *itemBegin) {
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? counts[i].all : 0;
        const unsigned long long excl = blockExclusive1024(v, warpSums, &carry);
        if (i < n) itemBegin[i] = excl;
    }
    if (threadIdx.x == 0) itemBegin[n] = carry;
}

// ------------------------------------------------------------------------------------------------
// Per-tile totals (written by the scan kernels with atomics) -> exclusive prefixes.
// tileTot / tileBase hold three arrays back to back: all matches, canonical, terminal non-canonical.
// ------------------------------------------------------------------------------------------------
// The totals, the overflow cursor and the redo counters are also published to `hostOut` (mapped
// pinned host memory: [0] ovCursor, [1..3] totals, [4] redo count | status << 32), so that the host
// needs no copy-engine transfer (which would queue behind the window counters going home).
__global__ void __launch_bounds__(1024) tileScanKernel(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *tileBase,
                                                       const unsigned int *ovCursor, const unsigned int *redo,
                                                       unsigned long long *hostOut) {
    // one block per array; a thread takes 16 consecutive tiles per round
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    const uint32_t k = blockIdx.x;
    if (threadIdx.x == 0) {
        carry = 0;
        if (k == 0) {
            hostOut[0] = *reinterpret_cast<const unsigned long long *>(ovCursor);
            hostOut[4] = redo ? (static_cast<unsigned long long>(redo[0]) | (static_cast<unsigned long long>(redo[1]) << 32)) : 0ull;
        }
    }
    __syncthreads();
    const uint32_t *in = tileTot + static_cast<size_t>(k) * nTiles;
    unsigned long long *out = tileBase + static_cast<size_t>(k) * (nTiles + 1);
    constexpr int kPer = 16;                                      // consecutive tiles per thread and round
    for (uint32_t base = 0; base < nTiles; base += 1024 * kPer) {
        const uint32_t i = base + threadIdx.x * kPer;
        uint32_t v[kPer];
        unsigned long long sum = 0;
#pragma unroll
        for (int j = 0; j < kPer; ++j) { v[j] = i + j < nTiles ? in[i + j] : 0u; sum += v[j]; }
        unsigned long long excl = blockExclusive1024(sum, warpSums, &carry);
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            if (i + j < nTiles) out[i + j] = excl;
            excl += v[j];
        }
    }
    if (threadIdx.x == 0) {
        out[nTiles] = carry;
        hostOut[1 + k] = carry;
        __threadfence_system();
    }
}

// ------------------------------------------------------------------------------------------------
// Slot arena -> ordered match list.  One block per tile, one warp per row slot at a time; fully
// coalesced 8-byte copies (all window / item accounting was done by the scan kernels).
// ------------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;

__global__ void __launch_bounds__(kGatherThreads) gatherKernel(const GatherParams g) {
    const uint32_t tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ uint32_t sOff[3][kTileRows + 1];
    if (warp < 3) {
        // exclusive prefix over the 32 row slots of this tile (warp k handles array k)
        const uint4 meta = g.slotMeta[static_cast<size_t>(tile) * kTileRows + lane];
        const uint32_t v = warp == 0 ? meta.x : warp == 1 ? meta.z : meta.w;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += u;
        }
        sOff[warp][lane] = incl - v;
        if (lane == 31) sOff[warp][kTileRows] = incl;
    }
    __syncthreads();
    if (sOff[0][kTileRows] == 0) return;
    const unsigned long long base = g.tileBase[tile];
    const unsigned long long baseCan = g.tileBase[static_cast<size_t>(g.nTiles) + 1 + tile];
    const unsigned long long baseNct = g.tileBase[2 * (static_cast<size_t>(g.nTiles) + 1) + tile];
    const uint32_t laneLt = (1u << lane) - 1u;

    // A warp owns the row slots warp, warp + 8, warp + 16, warp + 24.  Common case (each of them holds at most 64
    // records, all inside the slot): the eight loads are issued back to back before anything is stored, so that
    // eight requests per warp are in flight instead of one — the copy is latency-bound otherwise.
    constexpr int kPerWarp = kTileRows / (kGatherThreads / 32);
    uint32_t cntR[kPerWarp];
    bool simple = true;
#pragma unroll
    for (int j = 0; j < kPerWarp; ++j) {
        const int s = warp + j * (kGatherThreads / 32);
        cntR[j] = sOff[0][s + 1] - sOff[0][s];
        simple = simple && cntR[j] <= 64u && cntR[j] <= g.slotCap;
    }
    if (simple) {
        uint2 rec[kPerWarp][2];
#pragma unroll
        for (int j = 0; j < kPerWarp; ++j) {
            const int s = warp + j * (kGatherThreads / 32);
            const uint2 *src = reinterpret_cast<const uint2 *>(g.slots) + (static_cast<size_t>(tile) * kTileRows + s) * g.slotCap;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                rec[j][h] = static_cast<uint32_t>(lane + 32 * h) < cntR[j] ? __ldg(src + lane + 32 * h) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int j = 0; j < kPerWarp; ++j) {
            const int s = warp + j * (kGatherThreads / 32);
            const bool anySel = sOff[1][s + 1] != sOff[1][s] || sOff[2][s + 1] != sOff[2][s];     // warp-uniform
            unsigned long long oc = baseCan + sOff[1][s], on = baseNct + sOff[2][s];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (static_cast<uint32_t>(32 * h) >= cntR[j]) break;                              // warp-uniform
                const uint32_t k = static_cast<uint32_t>(lane + 32 * h);
                const bool have = k < cntR[j];
                const unsigned long long o = base + sOff[0][s] + k;
                if (have && o < g.matchCap) reinterpret_cast<uint2 *>(g.matches)[o] = make_uint2(rec[j][h].x, rec[j][h].y & 0x00FFFFFFu);
                if (!anySel) continue;
                const uint32_t flags = (rec[j][h].y >> 16) & 0xFFu;
                const bool isC = have && (flags & kMatchCan);
                const bool isN = have && (flags & (kMatchCan | kMatchTerm)) == kMatchTerm;
                const uint32_t bc = __ballot_sync(0xFFFFFFFFu, isC), bn = __ballot_sync(0xFFFFFFFFu, isN);
                if (isC && oc + __popc(bc & laneLt) < g.selCap) g.canIdx[oc + __popc(bc & laneLt)] = static_cast<uint32_t>(o);
                if (isN && on + __popc(bn & laneLt) < g.selCap) g.nctIdx[on + __popc(bn & laneLt)] = static_cast<uint32_t>(o);
                oc += __popc(bc);
                on += __popc(bn);
            }
        }
        return;
    }

    for (int s = warp; s < kTileRows; s += kGatherThreads / 32) {
        const uint32_t cnt = sOff[0][s + 1] - sOff[0][s];
        if (cnt == 0) continue;
        const size_t slotIdx = static_cast<size_t>(tile) * kTileRows + s;
        const uint2 *src = reinterpret_cast<const uint2 *>(g.slots) + slotIdx * g.slotCap;
        const unsigned long long ov = g.slotMeta[slotIdx].y;
        const unsigned long long o = base + sOff[0][s];
        unsigned long long oc = baseCan + sOff[1][s], on = baseNct + sOff[2][s];
        const bool anySel = sOff[1][s + 1] != sOff[1][s] || sOff[2][s + 1] != sOff[2][s];     // warp-uniform
        for (uint32_t k0 = 0; k0 < cnt; k0 += 32) {
            const uint32_t k = k0 + lane;
            uint2 rec = make_uint2(0u, 0u);
            bool have = k < cnt;
            if (have) {
                if (k < g.slotCap) rec = src[k];
                else if (ov + (k - g.slotCap) < g.ovCap) rec = reinterpret_cast<const uint2 *>(g.ovArena)[ov + (k - g.slotCap)];
                else have = false;                              // overflow arena exhausted: the host reruns the scan
            }
            if (have && o + k < g.matchCap) reinterpret_cast<uint2 *>(g.matches)[o + k] = make_uint2(rec.x, rec.y & 0x00FFFFFFu);
            if (!anySel) continue;
            // index lists of the -m selections, in genome order
            const uint32_t flags = (rec.y >> 16) & 0xFFu;
            const bool isC = have && (flags & kMatchCan);
            const bool isN = have && (flags & (kMatchCan | kMatchTerm)) == kMatchTerm;
            const uint32_t bc = __ballot_sync(0xFFFFFFFFu, isC), bn = __ballot_sync(0xFFFFFFFFu, isN);
            if (isC && oc + __popc(bc & laneLt) < g.selCap) g.canIdx[oc + __popc(bc & laneLt)] = static_cast<uint32_t>(o + k);
            if (isN && on + __popc(bn & laneLt) < g.selCap) g.nctIdx[on + __popc(bn & laneLt)] = static_cast<uint32_t>(o + k);
            oc += __popc(bc);
            on += __popc(bn);
        }
    }
}

cudaError_t launchTileScan(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *tileBase, const unsigned int *ovCursor,
                           const unsigned int *redo, unsigned long long *hostOut, cudaStream_t s) {
    tileScanKernel<<<3, 1024, 0, s>>>(tileTot, nTiles, tileBase, ovCursor, redo, hostOut);
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
    __device__ void begin(uint32_t pos, uint32_t len, uint32_t flags) {
        const uint32_t f = flags & kMatchFwd ? 1u : 0u, c = flags & kMatchCan ? 1u : 0u;
        start = pos; end = pos + len; prev = pos; n = 1; nFwd = f; nCan = c;
        cov = len; fwdCov = f * len; canCov = c * len;
    }
    __device__ void absorb(uint32_t pos, uint32_t len, uint32_t flags) {
        const uint32_t f = flags & kMatchFwd ? 1u : 0u, c = flags & kMatchCan ? 1u : 0u;
        ++n; nFwd += f; nCan += c; cov += len; fwdCov += f * len; canCov += c * len; prev = pos;
    }
};

__device__ void emitBlock(const BlockParams &p, const DevBlock &b) {
    const unsigned int slot = atomicAdd(p.blockCount, 1u);
    if (slot < p.blockCap) p.blocks[slot] = b;
}

}  // namespace

// One warp per (item, arm).  arm 0 = p (forward-strand matches from the start), arm 1 = q
// (reverse-strand matches from the end).  The warp loads 32 records at a time (coalesced) and
// all lanes run the same sequential state machine on the records of the wanted strand, handed
// round by shuffles; the walk stops at the first run start outside the terminal zone.
__global__ void __launch_bounds__(128) terminalKernel(const BlockParams p) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= 2u * p.nItems) return;
    const uint32_t item = gw >> 1;
    const bool fromStart = (gw & 1u) == 0;
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
        const uint32_t want = fromStart ? kMatchFwd : 0;
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
                if (lane == 0) emitBlock(p, cur);
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
        bool inRun = false, done = false;
        const uint2 *recs = reinterpret_cast<const uint2 *>(p.matches);
        // lane l holds the l-th next record in walk order; the next 32 are requested before the current 32 are walked
        auto fetch = [&](long long off, uint2 &rec, bool &valid) {
            const long long i = fromStart ? begin + off + lane : end - 1 - off - lane;
            valid = off < end - begin && (fromStart ? i < end : i >= begin);
            rec = valid ? __ldg(recs + i) : make_uint2(0u, 0u);
        };
        uint2 rec, recNext;
        bool valid, validNext;
        fetch(0, rec, valid);
        for (long long off = 0; off < end - begin && !done; off += 32, rec = recNext, valid = validNext) {
            fetch(off + 32, recNext, validNext);
            uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid && ((rec.y >> 16) & kMatchFwd) == want);
            if (inRun && mask) {
                // Inside a telomere the whole group of 32 continues the current run (every start-to-start gap
                // <= maxMatchDist; the zone test only applies to run starts, src/teloscope.cpp:98-121): absorb it
                // with a few warp reductions instead of one sequential step per match.
                const uint32_t below = mask & ((1u << lane) - 1u);
                uint32_t prevPos = __shfl_sync(0xFFFFFFFFu, rec.x, below ? 31 - __clz(below) : 0);
                if (!below) prevPos = run.prev;
                const bool mine = (mask >> lane) & 1u;
                const uint32_t gapL = fromStart ? rec.x - prevPos : prevPos - rec.x;
                if (!__any_sync(0xFFFFFFFFu, mine && gapL > c.maxMatchDist)) {
                    const uint32_t len = rec.y & 0xFFFFu;
                    const bool can = mine && ((rec.y >> 16) & kMatchCan);
                    const uint32_t cntM = __popc(mask), lenSum = __reduce_add_sync(0xFFFFFFFFu, mine ? len : 0u);
                    const uint32_t canLen = __reduce_add_sync(0xFFFFFFFFu, can ? len : 0u);
                    const int last = 31 - __clz(mask);
                    const uint32_t lastPos = __shfl_sync(0xFFFFFFFFu, rec.x, last), lastLen = __shfl_sync(0xFFFFFFFFu, len, last);
                    run.n += cntM; run.nFwd += want ? cntM : 0u; run.nCan += __popc(__ballot_sync(0xFFFFFFFFu, can));
                    run.cov += lenSum; run.fwdCov += want ? lenSum : 0u; run.canCov += canLen; run.prev = lastPos;
                    if (fromStart) run.end = lastPos + lastLen; else run.start = lastPos;
                    continue;
                }
            }
            while (mask && !done) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                const uint32_t pos = __shfl_sync(0xFFFFFFFFu, rec.x, b), hi = __shfl_sync(0xFFFFFFFFu, rec.y, b);
                const uint32_t len = hi & 0xFFFFu, flags = (hi >> 16) & 0xFFu;
                if (!inRun) {
                    if (!inZone(pos)) { done = true; break; }
                    run.begin(pos, len, flags);
                    inRun = true;
                    continue;
                }
                const uint32_t gap = fromStart ? pos - run.prev : run.prev - pos;
                if (gap <= c.maxMatchDist) {
                    if (fromStart) run.end = pos + len; else run.start = pos;
                    run.absorb(pos, len, flags);
                } else {
                    feedSub(run);
                    inRun = false;
                    if (!inZone(pos)) { done = true; break; }
                    run.begin(pos, len, flags);
                    inRun = true;
                }
            }
        }
        if (inRun) feedSub(run);
        if (haveCur) finishCur();
    }
    if (lane == 0) {
        if (fromStart) p.fwdBoundary[item] = boundary; else p.revBoundary[item] = boundary;
        if (readMode && emitted) p.pass[item] = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Interstitial blocks.  A block needs canonicalCount >= 4 (src/teloscope.cpp:186), so runs are found
// from the canonical matches: a seed thread per canonical match decides whether it is the first
// canonical match of its run and whether the run reaches 4 of them; a warp then walks each such run.
// Runs start afresh at the first match at or after fwdBoundary and stop before revBoundary (:188-191,235).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) itsSeedKernel(const BlockParams p, const uint32_t *canIdx, unsigned long long nCan,
                                                     unsigned long long *runStarts, unsigned int *runCount, uint32_t runCap) {
    const unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= nCan) return;
    const DevConfig &c = p.cfg;
    const unsigned long long idx = canIdx[j];
    uint32_t lo = 0, hi = p.nItems;                               // item: last i with itemBegin[i] <= idx
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.itemBegin[mid] <= idx) lo = mid + 1; else hi = mid;
    }
    const uint32_t item = lo - 1;
    const unsigned long long itemBegin = p.itemBegin[item], itemEnd = p.itemBegin[item + 1];
    const uint32_t fwdB = p.fwdBoundary[item], revB = p.revBoundary[item];
    if (!(fwdB < revB) || itemEnd - itemBegin < 2) return;        // src/teloscope.cpp:653
    const uint2 *recs = reinterpret_cast<const uint2 *>(p.matches);
    const uint32_t pos0 = recs[idx].x;
    if (pos0 < fwdB || pos0 >= revB) return;
    // walk left to the run start; give up if an earlier canonical match of the same run exists
    unsigned long long i = idx;
    uint32_t pos = pos0;
    while (i > itemBegin) {
        const uint2 pr = recs[i - 1];
        if (pr.x < fwdB || pos - pr.x > c.maxMatchDist) break;
        if ((pr.y >> 16) & kMatchCan) return;
        --i;
        pos = pr.x;
    }
    // walk right until the run holds 4 canonical matches (or ends)
    uint32_t nCanRun = 1;
    unsigned long long k = idx;
    pos = pos0;
    while (nCanRun < 4 && k + 1 < itemEnd) {
        const uint2 nx = recs[k + 1];
        if (nx.x >= revB || nx.x - pos > c.maxMatchDist) break;
        if ((nx.y >> 16) & kMatchCan) ++nCanRun;
        ++k;
        pos = nx.x;
    }
    if (nCanRun < 4) return;
    const unsigned int slot = atomicAdd(runCount, 1u);
    if (slot < runCap) runStarts[slot] = (static_cast<unsigned long long>(item) << 32) | i;   // i < 2^32
}

__global__ void __launch_bounds__(128) itsRunKernel(const BlockParams p, const unsigned long long *runStarts,
                                                    const unsigned int *runCount, uint32_t runCap) {
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t nRuns = min(*runCount, runCap);
    if (gw >= nRuns) return;
    const DevConfig &c = p.cfg;
    const unsigned long long e = runStarts[gw];
    const uint32_t item = static_cast<uint32_t>(e >> 32);
    const unsigned long long first = e & 0xFFFFFFFFull;
    const unsigned long long itemEnd = p.itemBegin[item + 1];
    const uint32_t revB = p.revBoundary[item];
    const uint2 *recs = reinterpret_cast<const uint2 *>(p.matches);
    uint32_t n = 0, nFwd = 0, nCan = 0, cov = 0, fwdCov = 0, canCov = 0, start = 0, end = 0;
    uint32_t prevPos = 0;
    bool open = true;
    for (unsigned long long base = first; open && base < itemEnd; base += 32) {
        const unsigned long long i = base + lane;
        const bool valid = i < itemEnd;
        const uint2 rec = valid ? recs[i] : make_uint2(0u, 0u);
        uint32_t before = __shfl_up_sync(0xFFFFFFFFu, rec.x, 1);
        if (lane == 0) before = prevPos;
        // the run's first record always belongs; the others need pos < revB and a start gap <= maxMatchDist
        const bool isFirst = base == first && lane == 0;
        const bool ok = valid && (isFirst || (rec.x < revB && rec.x - before <= c.maxMatchDist));
        const uint32_t bad = ~__ballot_sync(0xFFFFFFFFu, ok);
        const uint32_t take = bad ? (1u << (__ffs(bad) - 1)) - 1u : 0xFFFFFFFFu;     // lanes before the first break
        const bool mine = (take >> lane) & 1u;
        const uint32_t len = rec.y & 0xFFFFu, flags = (rec.y >> 16) & 0xFFu;
        const uint32_t f = (mine && (flags & kMatchFwd)) ? 1u : 0u, cn = (mine && (flags & kMatchCan)) ? 1u : 0u;
        const uint32_t cntTake = __popc(take);
        n += cntTake;
        nFwd += __popc(__ballot_sync(0xFFFFFFFFu, f != 0));
        nCan += __popc(__ballot_sync(0xFFFFFFFFu, cn != 0));
        cov += __reduce_add_sync(0xFFFFFFFFu, mine ? len : 0u);
        fwdCov += __reduce_add_sync(0xFFFFFFFFu, f ? len : 0u);
        canCov += __reduce_add_sync(0xFFFFFFFFu, cn ? len : 0u);
        if (base == first) start = __shfl_sync(0xFFFFFFFFu, rec.x, 0);
        if (cntTake) {
            const int last = static_cast<int>(cntTake) - 1;
            prevPos = __shfl_sync(0xFFFFFFFFu, rec.x, last);
            end = prevPos + __shfl_sync(0xFFFFFFFFu, len, last);
        }
        open = bad == 0;
    }
    if (lane != 0) return;
    const uint32_t len = end - start;
    // computeBlockLabel, include/teloscope.h:217-222 (float32)
    const float ratio = __fdiv_rn(__fmul_rn(__uint2float_rn(nFwd), 100.0f), __uint2float_rn(n));
    const char label = ratio > 66.6f ? 'p' : (ratio < 33.3f ? 'q' : 'b');
    if (len >= c.itsMinLen && nCan >= 4 && !(label == 'b' && nFwd < 2 && (n - nFwd) < 2)) {
        DevBlock b{};
        b.item = item; b.order = 0; b.start = start; b.len = len; b.counts = n; b.fwdCount = nFwd;
        b.canCount = nCan; b.totalCov = cov; b.fwdCov = fwdCov; b.canCov = canCov;
        b.validOr = 1; b.label = label; b.kind = 1;
        emitBlock(p, b);
    }
}

// ------------------------------------------------------------------------------------------------
// -m lists: final host-visible records (same layout as tls_match in include/teloscope_b200.h)
// ------------------------------------------------------------------------------------------------
struct OutMatch {
    unsigned long long pos;
    uint16_t len;
    uint8_t isForward, isCanonical;
    uint32_t pad;
};
static_assert(sizeof(OutMatch) == 16, "tls_match layout");

// compact form (tls_match8): the position stays relative to the item, so no item lookup is needed
__global__ void __launch_bounds__(256) selectFinalize8Kernel(const DevMatch *m, const uint32_t *idxList, unsigned long long n, uint2 *out) {
    const unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const DevMatch x = m[idxList[j]];
    out[j] = make_uint2(x.pos, static_cast<uint32_t>(x.len) | ((x.flags & kMatchFwd) ? 0x10000u : 0u) | ((x.flags & kMatchCan) ? 0x1000000u : 0u));
}

// tls_window (8 x u32) -> tls_window16 (8 x u16); tls_create has made sure that no counter can exceed 16 bits
__global__ void __launch_bounds__(256) packWindowsKernel(const uint4 *in, uint4 *out, unsigned long long n) {
    const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a = in[2 * i], b = in[2 * i + 1];
    out[i] = make_uint4(a.x | (a.y << 16), a.z | (a.w << 16), b.x | (b.y << 16), b.z | (b.w << 16));
}

__global__ void __launch_bounds__(256) selectFinalizeKernel(const DevMatch *m, const uint32_t *idxList, unsigned long long n,
                                                            const unsigned long long *itemBegin, const unsigned long long *absPos,
                                                            uint32_t nItems, OutMatch *out) {
    const unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned long long idx = idxList[j];
    uint32_t lo = 0, hi = nItems;                                  // item: last i with itemBegin[i] <= idx
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (itemBegin[mid] <= idx) lo = mid + 1; else hi = mid;
    }
    const DevMatch x = m[idx];
    OutMatch o;
    o.pos = absPos[lo - 1] + x.pos; o.len = x.len;
    o.isForward = x.flags & kMatchFwd ? 1 : 0; o.isCanonical = x.flags & kMatchCan ? 1 : 0; o.pad = 0;
    out[j] = o;
}

// ------------------------------------------------------------------------------------------------
// Window table for overlapping windows (step < window; src/teloscope.cpp:409-434,485,512-529,600-637).
// Window k of a contig starts at k * step and has min(window, n - k * step) bases; its counters are
// what the reference builds from two pieces: the carry of window k - 1 (everything found at i >= step
// there: start inside [k * step, ...) and end inside window k - 1) and its own part (match END index
// j >= overlap).  For step >= L and overlap >= L (L = the pattern length; tls_create refuses anything
// else) the two pieces add up to: every match that lies entirely inside the window, and every A/C/G/T
// base of the window; and across windows every match of the contig is reported exactly once (by the
// window its end falls into), so the scan kernels run without window truncation and this kernel builds
// the table from the sequence and the ordered match list.  One warp per window.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) overlapWindowsKernel(const uint8_t *seq, const DevRegion *regions, uint32_t nRegions,
                                                            const DevMatch *matches, const unsigned long long *itemBegin,
                                                            uint32_t nWindows, uint32_t W, uint32_t S, uint32_t countAcgt,
                                                            uint4 *windows) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (g >= nWindows) return;
    uint32_t lo = 0, hi = nRegions;                                  // last region with winBase <= g
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (regions[mid].winBase <= g) lo = mid; else hi = mid;
    }
    const DevRegion r = regions[lo];
    const uint32_t ws = (g - r.winBase) * S;                         // < r.len
    const uint32_t cw = min(W, r.len - ws), wEnd = ws + cw;
    uint32_t a = 0, c = 0, gg = 0, t = 0;
    if (countAcgt) {
        const uint8_t *row = seq + r.bufStart + ws;
        for (uint32_t i = lane; i < cw; i += 32) {
            const uint32_t ch = row[i] & 0xDFu;
            a += ch == 'A'; c += ch == 'C'; gg += ch == 'G'; t += ch == 'T';
        }
    }
    // first record of the item with pos >= ws
    unsigned long long b = itemBegin[r.item], e = itemBegin[r.item + 1];
    const unsigned long long itemEnd = e;
    while (b < e) {
        const unsigned long long mid = (b + e) >> 1;
        if (matches[mid].pos < ws) b = mid + 1; else e = mid;
    }
    uint32_t can = 0, non = 0, fwd = 0, rev = 0;
    for (unsigned long long base = b; base < itemEnd; base += 32) {
        const unsigned long long i = base + lane;
        bool inside = false;
        if (i < itemEnd) {
            const DevMatch m = matches[i];
            inside = m.pos < wEnd;
            if (inside && m.pos + m.len <= wEnd) {                   // entirely inside the window
                if (m.flags & kMatchCan) can += m.len; else non += m.len;
                if (m.flags & kMatchFwd) fwd += m.len; else rev += m.len;
            }
        }
        if (!__all_sync(0xFFFFFFFFu, inside)) break;
    }
    a = __reduce_add_sync(0xFFFFFFFFu, a); c = __reduce_add_sync(0xFFFFFFFFu, c);
    gg = __reduce_add_sync(0xFFFFFFFFu, gg); t = __reduce_add_sync(0xFFFFFFFFu, t);
    can = __reduce_add_sync(0xFFFFFFFFu, can); non = __reduce_add_sync(0xFFFFFFFFu, non);
    fwd = __reduce_add_sync(0xFFFFFFFFu, fwd); rev = __reduce_add_sync(0xFFFFFFFFu, rev);
    if (lane == 0) {
        windows[2 * static_cast<size_t>(g)] = make_uint4(a, c, gg, t);
        windows[2 * static_cast<size_t>(g) + 1] = make_uint4(can, non, fwd, rev);
    }
}

cudaError_t launchOverlapWindows(const uint8_t *seq, const DevRegion *regions, uint32_t nRegions, const DevMatch *matches,
                                 const unsigned long long *itemBegin, uint32_t nWindows, const DevConfig &cfg, uint32_t *windows,
                                 cudaStream_t s) {
    if (nWindows == 0 || nRegions == 0) return cudaSuccess;
    const unsigned long long threads = 32ull * nWindows;
    overlapWindowsKernel<<<static_cast<unsigned int>((threads + 255) / 256), 256, 0, s>>>(
        seq, regions, nRegions, matches, itemBegin, nWindows, cfg.windowSize, cfg.step, cfg.countAcgt, reinterpret_cast<uint4 *>(windows));
    return cudaGetLastError();
}
cudaError_t launchSelectFinalize8(const DevMatch *m, const uint32_t *idxList, unsigned long long n, void *out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    selectFinalize8Kernel<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, s>>>(m, idxList, n, static_cast<uint2 *>(out));
    return cudaGetLastError();
}
cudaError_t launchPackWindows(const uint32_t *windows, void *windows16, unsigned long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    packWindowsKernel<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4 *>(windows),
                                                                              static_cast<uint4 *>(windows16), n);
    return cudaGetLastError();
}
cudaError_t launchItemOffsets(const DevItemCounts *counts, uint32_t n, unsigned long long *itemBegin, cudaStream_t s) {
    itemOffsetsKernel<<<1, 1024, 0, s>>>(counts, n, itemBegin);
    return cudaGetLastError();
}
cudaError_t launchTerminal(const BlockParams &p, cudaStream_t s) {
    const unsigned long long threads = 64ull * p.nItems;              // one warp per (item, arm)
    if (threads == 0) return cudaSuccess;
    terminalKernel<<<static_cast<unsigned int>((threads + 127) / 128), 128, 0, s>>>(p);
    return cudaGetLastError();
}
cudaError_t launchIts(const BlockParams &p, const uint32_t *canIdx, unsigned long long nCan, unsigned long long *runStarts,
                      unsigned int *runCount, uint32_t runCap, cudaStream_t s) {
    if (nCan == 0) return cudaSuccess;
    itsSeedKernel<<<static_cast<unsigned int>((nCan + 255) / 256), 256, 0, s>>>(p, canIdx, nCan, runStarts, runCount, runCap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the number of runs is only known on the device: enough warps for the capacity, surplus warps exit
    const unsigned long long warps = std::min<unsigned long long>(runCap, nCan / 4 + 1);
    itsRunKernel<<<static_cast<unsigned int>((warps * 32 + 127) / 128), 128, 0, s>>>(p, runStarts, runCount, runCap);
    return cudaGetLastError();
}
cudaError_t launchSelectFinalize(const DevMatch *m, const uint32_t *idxList, unsigned long long n,
                                 const unsigned long long *itemBegin, const unsigned long long *absPos, uint32_t nItems,
                                 void *out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    selectFinalizeKernel<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, s>>>(m, idxList, n, itemBegin, absPos, nItems,
                                                                                   static_cast<OutMatch *>(out));
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
