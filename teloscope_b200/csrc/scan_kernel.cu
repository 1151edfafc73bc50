// General telomeric motif scan kernel for sm_100a (any mode, any pattern set, any window size).
// Tiles of a genome-mode batch that lie inside one contig and use one short pattern length go to
// the lean kernel in scan_fast.cu instead; this kernel takes every other tile plus the rows the
// lean kernel handed back (rows holding a non-ACGT byte).
//
// One persistent CTA per SM walks its work items (<= 49136 bases of the staged sequence buffer), tile
// index = blockIdx.x + k * gridDim.x.  Nothing in the kernel waits on another CTA.  Per tile:
//   1. the raw ASCII bytes are staged into shared memory with cp.async.bulk (TMA bulk copy,
//      mbarrier completion), double buffered; the next tile's descriptor and regions are
//      prefetched into registers one iteration ahead so no global latency sits on the tile path;
//   2. every lane packs its 48-base span to 2-bit codes in registers (SWAR multiply-gather) and
//      probes a 64 KiB shared-memory table with the 8-mer at every NPOS-th offset; one byte of
//      the table answers NPOS start positions at once (replaces the reference's Trie walk,
//      include/teloscope.h:15-86, src/teloscope.cpp:437-445 and :546-554);
//   3. candidate positions are compacted per warp (scan + scatter into a shared-memory queue) and
//      resolved 32 at a time: exact ASCII validity, window / region truncation
//      (src/teloscope.cpp:439,:549), orientation + canonical flags, per-window covered-bp and
//      per-item counters;
//   4. A/C/G/T window counts come from popcounts on the packed codes (src/teloscope.cpp:418-435);
//   5. match records go, in genome order, to the (tile, warp) slot of a global arena; a later
//      scan + gather pass (block_kernels.cu) compacts the slots into one ordered list.
// No tensor cores: this is byte/integer streaming work bound by instruction issue and HBM.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_common.cuh"
#include "kernels.hpp"

namespace tls {

namespace {

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smemAddr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smemAddr(dst)),
        "l"(src), "r"(bytes), "r"(smemAddr(bar))
        : "memory");
}

// ------------------------------------------------------------------------------------------------
// SWAR helpers on 4 ASCII bytes per 32-bit word
// ------------------------------------------------------------------------------------------------

// 2-bit codes of 4 bases -> 8 contiguous bits.  (c>>1)&3 maps A,C,T,G (either case) to 0,1,2,3;
// the multiply gathers bits {1,2} of each byte into bits 24..31 (no two partial products share
// a bit, so there are no carries).
__device__ __forceinline__ uint32_t pack4(uint32_t w) { return ((w & 0x06060606u) * 0x00820820u) >> 24; }

__device__ __forceinline__ uint32_t pack16(const uint4 &v) {
    const uint32_t a = (v.x & 0x06060606u) * 0x00820820u;
    const uint32_t b = (v.y & 0x06060606u) * 0x00820820u;
    const uint32_t c = (v.z & 0x06060606u) * 0x00820820u;
    const uint32_t d = (v.w & 0x06060606u) * 0x00820820u;
    // bytes 3 of a,b,c,d -> one word (a lowest)
    return __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
}

// Bit 4 of a byte of the result is set iff that byte violates b0 != b4 or (b2 & ~b1) == b4; with
// b7=0, b6=1, b3=0 (checked separately on OR/AND aggregates) these relations single out ACGTacgt.
__device__ __forceinline__ uint32_t relBits(uint32_t w) {
    const uint32_t l2 = w << 2, l3 = w << 3, l4 = w << 4;
    return (~(l4 ^ w)) | ((l2 & ~l3) ^ w);
}
// Nonzero iff some byte of w is not one of ACGTacgt; flagged bytes carry a bit in 0xD8.
__device__ __forceinline__ uint32_t badBytes(uint32_t w) {
    return (relBits(w) & 0x10101010u) | ((w ^ 0x40404040u) & 0xC8C8C8C8u);
}
// Same predicate for 16 bytes at once (only "any bad byte", no positions).
__device__ __forceinline__ uint32_t badChunk(const uint4 &v) {
    const uint32_t rel = relBits(v.x) | relBits(v.y) | relBits(v.z) | relBits(v.w);
    const uint32_t anyOr = v.x | v.y | v.z | v.w, allAnd = v.x & v.y & v.z & v.w;
    return (rel & 0x10101010u) | (anyOr & 0x88888888u) | (~allAnd & 0x40404040u);
}

__device__ __forceinline__ bool isBase(uint32_t c) {
    c &= 0xDFu;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

__device__ __forceinline__ uint32_t evenMaskBelow(int n) {   // even bits 0,2,..,2(n-1); n clamped to [0,16]
    n = n < 0 ? 0 : n;
    return n >= 16 ? 0x55555555u : (((1u << (2 * n)) - 1u) & 0x55555555u);
}

// Per-tile facts every thread needs, written to shared memory by warp 0 before the tile starts.
struct TileMeta {
    uint64_t bufStart;
    uint32_t len;
    uint32_t firstRegion;
    uint32_t nCached;       // regions of this tile in shared memory
    uint32_t firstWin;      // window index of shared counter slot 0
    uint32_t single;        // 1: region 0 covers the whole tile
    uint32_t relBase;       // single: item-relative position of tile byte 0
};

struct Located {
    bool ok;
    uint32_t cacheIdx;      // index in the shared region cache or 0xFFFFFFFF
    uint32_t item, rel, itemLen, regionEndRel, winBase;
};

__device__ __forceinline__ Located locate(const ScanParams &p, const TileMeta &t, const DevRegion *sRegions, int q) {
    Located r;
    r.ok = false;
    r.cacheIdx = 0xFFFFFFFFu;
    if (q < 0 || q >= static_cast<int>(t.len)) return r;
    if (t.single) {
        const DevRegion &c = sRegions[0];
        r.ok = true; r.cacheIdx = 0; r.item = c.item; r.rel = t.relBase + static_cast<uint32_t>(q);
        r.itemLen = c.itemLen; r.regionEndRel = c.relStart + c.len; r.winBase = c.winBase;
        return r;
    }
    const uint64_t pos = t.bufStart + static_cast<uint32_t>(q);
    const DevRegion *reg = nullptr;
    for (uint32_t k = 0; k < t.nCached; ++k) {
        const DevRegion &c = sRegions[k];
        if (pos >= c.bufStart && pos < c.bufStart + c.len) { reg = &c; r.cacheIdx = k; break; }
        if (pos < c.bufStart) return r;                     // padding between regions
    }
    if (!reg) {
        if (t.nCached < kRegionCache) return r;
        // more regions than the cache holds: binary search the global table
        uint32_t lo = t.firstRegion + kRegionCache, hi = p.nRegions;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.regions[mid].bufStart <= pos) lo = mid + 1; else hi = mid;
        }
        if (lo == t.firstRegion + kRegionCache) return r;
        reg = &p.regions[lo - 1];
        if (pos >= reg->bufStart + reg->len) return r;
    }
    r.ok = true;
    r.item = reg->item;
    r.rel = reg->relStart + static_cast<uint32_t>(pos - reg->bufStart);
    r.itemLen = reg->itemLen;
    r.regionEndRel = reg->relStart + reg->len;
    r.winBase = reg->winBase;
    return r;
}

// Warp-level "add by key": one leader per distinct key among the active lanes receives the sums.
template <typename F>
__device__ __forceinline__ void warpAddByKey(uint32_t active, uint32_t key, uint32_t v0, uint32_t v1, F f) {
    const int lane = threadIdx.x & 31;
    while (active) {
        const int leader = __ffs(active) - 1;
        const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, key, leader);
        const uint32_t grp = __ballot_sync(0xFFFFFFFFu, key == k0) & active;
        const bool mine = (grp >> lane) & 1u;
        const uint32_t s0 = __reduce_add_sync(0xFFFFFFFFu, mine ? v0 : 0u);
        const uint32_t s1 = __reduce_add_sync(0xFFFFFFFFu, mine ? v1 : 0u);
        if (lane == leader) f(k0, s0, s1);
        active &= ~grp;
    }
}

__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

}  // namespace

// ================================================================================================
// The fused scan kernel
// ================================================================================================
template <int NPOS, bool COUNT>
__global__ void __launch_bounds__(kScanThreads, 1) scanKernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *sIndex = smem;
    uint8_t *sFlags = sIndex + 65536;
    uint8_t *sTile = sFlags + kFlagTabBytes;
    constexpr int kTileStride = kTileSmemBytes + kTileHalo;
    uint16_t *sQueueAll = reinterpret_cast<uint16_t *>(sTile + 2 * kTileStride);
    DevRegion *sRegionsAll = reinterpret_cast<DevRegion *>(sQueueAll + kScanWarps * kQueueCap);   // [2][cache]
    uint32_t *sWinAll = reinterpret_cast<uint32_t *>(sRegionsAll + 2 * kRegionCache);              // [2][slots][4]
    TileMeta *sMetaAll = reinterpret_cast<TileMeta *>(sWinAll + 2 * kWindowSlots * 4);             // [2]
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sMetaAll + 2);                                   // [2]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t laneLt = (1u << lane) - 1u;
    const DevConfig &cfg = p.cfg;
    uint16_t *sQueue = sQueueAll + warp * kQueueCap;

    // ---- work list: whole tiles, then single rows handed back by the lean kernel ---------------------
    const uint32_t nRedo = p.redoCount ? min(*p.redoCount, p.redoCap) : 0u;
    const uint32_t nWorkItems = p.nWork + nRedo;
    const uint32_t G = gridDim.x;
    uint32_t work = blockIdx.x;
    if (work >= nWorkItems) return;
    auto workTile = [&](uint32_t i, uint32_t &rows) -> uint32_t {
        if (i < p.nWork) { rows = 0xFFFFFFFFu; return p.workTiles ? p.workTiles[i] : i; }
        const uint32_t e = p.redoList[i - p.nWork];
        rows = 1u << (e & 31u);
        return e >> 5;
    };

    // ---- one-time setup: tables to shared memory, barriers --------------------------------------
    for (int i = tid; i < 65536 / 16; i += kScanThreads)
        reinterpret_cast<uint4 *>(sIndex)[i] = reinterpret_cast<const uint4 *>(p.index8)[i];
    for (uint32_t i = tid; i < cfg.flagTabBytes / 16; i += kScanThreads)
        reinterpret_cast<uint4 *>(sFlags)[i] = reinterpret_cast<const uint4 *>(p.flagTabs)[i];
    if (tid == 0) {
        mbarInit(&sBar[0], 1);
        mbarInit(&sBar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issueLoad = [&](const DevTile &t, int buf) {   // one thread
        const uint64_t a = t.bufStart & ~15ull;
        const uint32_t bytes = (static_cast<uint32_t>(t.bufStart - a) + t.len + kTileHalo + 15u) & ~15u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&sBar[buf], bytes);
        uint8_t *dst = sTile + buf * kTileStride;
        for (uint32_t off = 0; off < bytes; off += 16384u)
            bulkLoad(dst + off, p.seq + a + off, min(16384u, bytes - off), &sBar[buf]);
    };
    auto loadRegion = [&](const DevTile &t) {           // threads of warp 0: candidate region `lane`
        const uint32_t idx = min(t.firstRegion + static_cast<uint32_t>(lane), p.nRegions - 1);
        return p.regions[idx];
    };

    // software pipeline registers: descriptors of work item k (cur), k+1 (next); regions of cur
    uint32_t rowsCur, rowsNext = 0, rowsNext2 = 0;
    uint32_t tileCur = workTile(work, rowsCur), tileNext = 0, tileNext2 = 0;
    DevTile tCur = p.tiles[tileCur];
    DevTile tNext{};
    if (work + G < nWorkItems) { tileNext = workTile(work + G, rowsNext); tNext = p.tiles[tileNext]; }
    DevRegion rCur{};
    if (warp == 0) rCur = loadRegion(tCur);
    if (tid == 0) issueLoad(tCur, 0);

    for (uint32_t iter = 0; work < nWorkItems; ++iter, work += G) {
        const int buf = iter & 1;
        const uint8_t *tile = sTile + buf * kTileStride;
        DevRegion *sRegions = sRegionsAll + buf * kRegionCache;
        uint32_t *sWin = sWinAll + buf * kWindowSlots * 4;
        const bool hasNext = work + G < nWorkItems;
        const uint32_t tileIdx = tileCur, rowMask = rowsCur;

        // ---- prefetch: data of item k+1, descriptor of k+2, regions of k+1 (all consumed later) ----
        if (tid == 0 && hasNext) issueLoad(tNext, buf ^ 1);
        DevTile tNext2{};
        if (work + 2 * G < nWorkItems) { tileNext2 = workTile(work + 2 * G, rowsNext2); tNext2 = p.tiles[tileNext2]; }
        DevRegion rNext{};
        if (warp == 0 && hasNext) rNext = loadRegion(tNext);

        // ---- publish this tile's metadata (registers -> shared) ----------------------------------------
        if (warp == 0) {
            const uint64_t tileEnd = tCur.bufStart + tCur.len;
            const bool in = tCur.firstRegion + lane < p.nRegions && rCur.bufStart < tileEnd;
            const uint32_t n = __popc(__ballot_sync(0xFFFFFFFFu, in));   // regions are sorted: a prefix
            if (in) sRegions[lane] = rCur;
            if (lane == 0) {
                TileMeta m;
                m.bufStart = tCur.bufStart; m.len = tCur.len; m.firstRegion = tCur.firstRegion; m.nCached = n;
                m.single = n >= 1 && rCur.bufStart <= tCur.bufStart && tileEnd <= rCur.bufStart + rCur.len;
                const uint64_t s = tCur.bufStart > rCur.bufStart ? tCur.bufStart : rCur.bufStart;
                const uint32_t relFirst = rCur.relStart + static_cast<uint32_t>(s - rCur.bufStart);
                m.relBase = rCur.relStart + static_cast<uint32_t>(tCur.bufStart - rCur.bufStart);   // valid if single
                m.firstWin = (cfg.mode == 0 && n > 0) ? rCur.winBase + divWindow(cfg, relFirst) : 0;
                sMetaAll[buf] = m;
            }
        }
        if (COUNT)
            for (int i = tid; i < kWindowSlots * 4; i += kScanThreads) sWin[i] = 0;
        __syncthreads();
        const TileMeta t = sMetaAll[buf];
        const uint32_t skew = static_cast<uint32_t>(t.bufStart & 15ull);
        const uint32_t regionEnd0 = sRegions[0].relStart + sRegions[0].len;   // single tiles
        const uint32_t winBase0 = sRegions[0].winBase;
        mbarWait(&sBar[buf], (iter >> 1) & 1u);

        auto addAcgt = [&](uint32_t win, uint32_t a, uint32_t c, uint32_t g, uint32_t tt) {   // one lane
            const uint32_t s = win - t.firstWin;
            uint32_t *w = s < kWindowSlots ? sWin + s * 4 : p.windows + static_cast<size_t>(win) * 8;
            if (a) atomicAdd(w + 0, a);
            if (c) atomicAdd(w + 1, c);
            if (g) atomicAdd(w + 2, g);
            if (tt) atomicAdd(w + 3, tt);
        };

        const uint32_t usedBytes = skew + t.len;
#pragma unroll 1
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
            const uint32_t row = warp * kRowsPerWarp + rr;
            if (!((rowMask >> row) & 1u)) continue;            // warp-uniform
            // ---- per-row state: every row owns one output slot ------------------------------------------
            uint32_t wMatches = 0;      // match records this row has produced
            uint32_t rowCan = 0, rowNct = 0;   // canonical / terminal non-canonical records among them
            uint32_t ovBase = kNoOverflow;
            const size_t slotIdx = static_cast<size_t>(tileIdx) * kTileRows + row;
            DevMatch *slot = p.slots + slotIdx * p.slotCap;
            uint32_t candLo = 0, candHi = 0;

            // ---- stage 2: verify up to 32 queued candidates and emit match records ---------------------
            auto resolveBatch = [&](uint32_t from, uint32_t count) {
                uint32_t hitMask = 0, flagBits = 0, rel = 0, termBit = 0;
                if (lane < static_cast<int>(count)) {
                    const uint32_t sb = sQueue[from + lane];
                    const uint32_t q = sb - skew;                       // tile position (wraps if before the tile)
                    bool ok = q < t.len;
                    uint32_t limit = 0, item = 0, itemLen = 0, winBaseLoc = 0;
                    if (t.single) {
                        rel = t.relBase + q;
                        limit = regionEnd0;
                    } else if (ok) {
                        const Located loc = locate(p, t, sRegions, static_cast<int>(q));
                        ok = loc.ok;
                        rel = loc.rel; limit = loc.regionEndRel; item = loc.item; itemLen = loc.itemLen; winBaseLoc = loc.winBase;
                    }
                    if (ok) {
                        if (cfg.mode == 0 && !cfg.overlapWindows) {     // matches never cross the window end (src/teloscope.cpp:439)
                            const uint64_t wEnd = static_cast<uint64_t>(divWindow(cfg, rel) + 1) * cfg.windowSize;
                            if (wEnd < limit) limit = static_cast<uint32_t>(wEnd);
                        }
                        const uint32_t room = limit - rel;              // >= 1
                        if (room >= cfg.minLen) {
                            // 8 bytes at sb (unaligned) as two words
                            const uint32_t a = sb & ~3u, sh = (sb & 3u) * 8u;
                            const uint32_t w0 = *reinterpret_cast<const uint32_t *>(tile + a);
                            const uint32_t w1 = *reinterpret_cast<const uint32_t *>(tile + a + 4);
                            const uint32_t w2 = *reinterpret_cast<const uint32_t *>(tile + a + 8);
                            const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
                            uint64_t key = pack4(x0) | (pack4(x1) << 8);
                            const uint32_t b0 = badBytes(x0), b1 = badBytes(x1);
                            uint32_t nValid = b0 ? (__ffs(b0) - 1) >> 3 : (b1 ? 4 + ((__ffs(b1) - 1) >> 3) : 8);
                            if (nValid == 8 && cfg.maxLen > 8 && room > 8) {   // long patterns: byte loop
                                const uint32_t need = min(cfg.maxLen, room);
                                for (uint32_t j = 8; j < need; ++j) {
                                    const uint32_t c = tile[sb + j];
                                    if (!isBase(c)) break;
                                    key |= static_cast<uint64_t>((c >> 1) & 3u) << (2 * j);
                                    ++nValid;
                                }
                            }
                            const uint32_t usable = min(nValid, room);
                            for (uint32_t ci = 0; ci < cfg.nClasses; ++ci) {
                                const DevLenClass &lc = cfg.cls[ci];
                                if (lc.len > usable) break;
                                const uint64_t k = lc.len >= 32 ? key : (key & ((1ull << (2 * lc.len)) - 1ull));
                                uint32_t f = 0;
                                if (lc.direct) {
                                    f = sFlags[lc.flagOffset + static_cast<uint32_t>(k)];
                                } else {
                                    uint32_t lo = lc.sortedBegin, hi = lc.sortedBegin + lc.sortedCount;
                                    while (lo < hi) {
                                        const uint32_t mid = (lo + hi) >> 1;
                                        if (p.sortedKeys[mid] < k) lo = mid + 1; else hi = mid;
                                    }
                                    if (lo < lc.sortedBegin + lc.sortedCount && p.sortedKeys[lo] == k) f = p.sortedFlags[lo];
                                }
                                if (f & 1u) {
                                    hitMask |= 1u << ci;
                                    flagBits |= ((f >> 1) & 3u) << (2 * ci);
                                }
                            }
                        }
                    }
                    if (hitMask && cfg.mode == 0) {          // "terminal" status, src/teloscope.cpp:451-459
                        const uint32_t nIt = t.single ? sRegions[0].itemLen : itemLen;
                        if (cfg.overlapWindows) {            // the window-level shortcuts of :403-406 agree with the per-match test
                            const uint32_t termEnd = nIt > cfg.terminalLimit ? nIt - cfg.terminalLimit : 0u;
                            if (rel <= cfg.terminalLimit || rel >= termEnd) termBit = kMatchTerm;
                        } else if (isTerminalMatch(cfg, rel, divWindow(cfg, rel), nIt)) termBit = kMatchTerm;
                    }
                    if (hitMask) {
                        // window coverage (src/teloscope.cpp:470-510) and per-item counters of this position's matches
                        if (t.single) { item = sRegions[0].item; winBaseLoc = winBase0; }
                        DevItemCounts *c = p.itemCounts + item;
                        const uint32_t winIdx = cfg.mode == 0 ? divWindow(cfg, rel) : 0;
                        const bool term = termBit != 0;
                        for (uint32_t hm = hitMask; hm; hm &= hm - 1) {
                            const int ci = __ffs(hm) - 1;
                            const uint32_t fl = (flagBits >> (2 * ci)) & 3u, len = cfg.cls[ci].len;
                            atomicAdd(&c->all, 1u);
                            atomicAdd((fl & 1u) ? &c->fwd : &c->rev, 1u);
                            if (fl & 2u) atomicAdd(&c->can, 1u);
                            else if (term) atomicAdd(&c->nonCanTerm, 1u);
                            if (cfg.mode == 0 && !cfg.overlapWindows) {
                                uint32_t *w = p.windows + static_cast<size_t>(winBaseLoc + winIdx) * 8;
                                atomicAdd(w + ((fl & 2u) ? 4 : 5), len);
                                atomicAdd(w + ((fl & 1u) ? 6 : 7), len);
                            }
                        }
                    }
                }
                // ---- ordered emission into the row's slot ---------------------------------------------
                const uint32_t nHit = __popc(hitMask);
                const uint32_t have = __ballot_sync(0xFFFFFFFFu, nHit != 0);
                if (have == 0) return;
                uint32_t excl, total;
                if (cfg.nClasses == 1) {
                    excl = __popc(have & laneLt);
                    total = __popc(have);
                } else {
                    const uint32_t incl = warpInclusiveScan(nHit);
                    total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    excl = incl - nHit;
                }
                if (wMatches + total > p.slotCap && ovBase == kNoOverflow) {
                    if (lane == 0) ovBase = takeOverflowChunk(p.ovCursor, p.ovChunk);
                    ovBase = __shfl_sync(0xFFFFFFFFu, ovBase, 0);
                }
                uint32_t o = wMatches + excl, nCanLane = 0, nNctLane = 0;
                for (uint32_t hm = hitMask; hm; hm &= hm - 1, ++o) {
                    const int ci = __ffs(hm) - 1;
                    DevMatch m;
                    m.pos = rel;
                    m.len = static_cast<uint16_t>(cfg.cls[ci].len);
                    m.flags = static_cast<uint8_t>(((flagBits >> (2 * ci)) & 3u) | termBit);   // kMatchFwd | kMatchCan
                    m.region = kRegionAccounted;
                    if (m.flags & kMatchCan) ++nCanLane; else if (termBit) ++nNctLane;
                    if (o < p.slotCap) slot[o] = m;
                    else if (static_cast<uint64_t>(ovBase) + (o - p.slotCap) < p.ovCap) p.ovArena[ovBase + (o - p.slotCap)] = m;
                }
                wMatches += total;
                rowCan += __reduce_add_sync(0xFFFFFFFFu, nCanLane);
                rowNct += __reduce_add_sync(0xFFFFFFFFu, nNctLane);
            };

            const uint32_t spanByte = row * kRowBytes + lane * kSpanBytes;
            if (row * kRowBytes < usedBytes) {                   // warp-uniform
            const uint4 *src = reinterpret_cast<const uint4 *>(tile + spanByte);
            const uint4 c0 = src[0], c1 = src[1], c2 = src[2], c3 = src[3];
            uint32_t W[5];
            W[0] = pack16(c0); W[1] = pack16(c1); W[2] = pack16(c2); W[3] = pack16(c3); W[4] = 0;

            if (COUNT) {
                // ---- A/C/G/T counts of this span, split at a window boundary -------------------------
                const int q0 = static_cast<int>(spanByte) - static_cast<int>(skew);
                const int lo = q0 < 0 ? -q0 : 0;                                  // first span position inside the tile
                int hi = static_cast<int>(t.len) - q0; hi = hi > kSpanBytes ? kSpanBytes : hi;   // one past the last
                uint32_t winA = 0xFFFFFFFFu, winB = 0xFFFFFFFFu;
                uint32_t acA = 0, gtA = 0, acB = 0, gtB = 0;                      // packed A|C<<16, G|T<<16
                if (lo < hi) {
                    bool fast = false;
                    uint32_t relA = 0, wa = 0, winBase = winBase0;
                    int split = kSpanBytes;
                    const uint32_t bad = badChunk(c0) | badChunk(c1) | badChunk(c2);
                    if (bad == 0) {
                        if (t.single) {
                            relA = t.relBase + static_cast<uint32_t>(q0 + lo);
                            fast = true;
                        } else {
                            const Located a = locate(p, t, sRegions, q0 + lo);
                            const Located z = locate(p, t, sRegions, q0 + hi - 1);
                            fast = a.ok && z.ok && a.item == z.item && z.rel - a.rel == static_cast<uint32_t>(hi - 1 - lo);
                            relA = a.rel; winBase = a.winBase;
                        }
                    }
                    if (fast) {
                        wa = divWindow(cfg, relA);
                        const uint64_t toNext = static_cast<uint64_t>(wa + 1) * cfg.windowSize - relA;   // >= 1
                        if (toNext < static_cast<uint64_t>(hi - lo)) {
                            split = lo + static_cast<int>(toNext);
                            // a second boundary inside the span (window < 48 bases) -> slow path
                            if (static_cast<uint64_t>(hi - split) > cfg.windowSize) fast = false;
                        }
                    }
                    if (fast) {
                        winA = winBase + wa;
                        if (split < hi) winB = winA + 1;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const uint32_t w = W[j], h = w >> 1;
                            const uint32_t inc = evenMaskBelow(hi - 16 * j) & ~evenMaskBelow(lo - 16 * j);
                            const uint32_t mA = inc & evenMaskBelow(split - 16 * j);
                            const uint32_t xC = w & ~h, xG = w & h, xT = ~w & h;
                            const uint32_t nC = __popc(xC & mA), nG = __popc(xG & mA), nT = __popc(xT & mA);
                            acA += (__popc(mA) - nC - nG - nT) | (nC << 16);
                            gtA += nG | (nT << 16);
                            if (mA != inc) {
                                const uint32_t mB = inc & ~mA;
                                const uint32_t nC2 = __popc(xC & mB), nG2 = __popc(xG & mB), nT2 = __popc(xT & mB);
                                acB += (__popc(mB) - nC2 - nG2 - nT2) | (nC2 << 16);
                                gtB += nG2 | (nT2 << 16);
                            }
                        }
                    } else {
                        // slow path: region edge, tiny windows, or a non-ACGT byte in the span
                        for (int i = lo; i < hi; ++i) {
                            const uint32_t c = tile[spanByte + i];
                            if (!isBase(c)) continue;
                            const Located m = locate(p, t, sRegions, q0 + i);
                            if (!m.ok) continue;
                            const uint32_t code = (c >> 1) & 3u;                // A0 C1 T2 G3
                            addAcgt(m.winBase + divWindow(cfg, m.rel), code == 0, code == 1, code == 3, code == 2);
                        }
                    }
                }
                // ---- reduce across the warp: windows touched by a row are consecutive ------------------
                const uint32_t act = __ballot_sync(0xFFFFFFFFu, winA != 0xFFFFFFFFu);
                if (act) {
                    const uint32_t wFirst = __shfl_sync(0xFFFFFFFFu, winA, __ffs(act) - 1);
                    const uint32_t lastWin = winB != 0xFFFFFFFFu ? winB : winA;
                    const uint32_t wLast = __shfl_sync(0xFFFFFFFFu, lastWin, 31 - __clz(act));
                    const uint32_t groups = wLast - wFirst + 1;
                    if (groups <= 4) {
                        for (uint32_t g = 0; g < groups; ++g) {
                            const uint32_t wg = wFirst + g;
                            const uint32_t v0 = (winA == wg ? acA : 0u) + (winB == wg ? acB : 0u);
                            const uint32_t v1 = (winA == wg ? gtA : 0u) + (winB == wg ? gtB : 0u);
                            const uint32_t s0 = __reduce_add_sync(0xFFFFFFFFu, v0), s1 = __reduce_add_sync(0xFFFFFFFFu, v1);
                            if (lane == 0) addAcgt(wg, s0 & 0xFFFFu, s0 >> 16, s1 & 0xFFFFu, s1 >> 16);
                        }
                    } else {       // several items in one row: window ids are not consecutive
                        warpAddByKey(act, winA, acA, gtA, [&](uint32_t k, uint32_t s0, uint32_t s1) {
                            addAcgt(k, s0 & 0xFFFFu, s0 >> 16, s1 & 0xFFFFu, s1 >> 16);
                        });
                        const uint32_t hasB = __ballot_sync(0xFFFFFFFFu, winB != 0xFFFFFFFFu);
                        warpAddByKey(hasB, winB, acB, gtB, [&](uint32_t k, uint32_t s0, uint32_t s1) {
                            addAcgt(k, s0 & 0xFFFFu, s0 >> 16, s1 & 0xFFFFu, s1 >> 16);
                        });
                    }
                }
            }

            // ---- stage 1: one table byte per NPOS start positions ---------------------------------
#pragma unroll
            for (int tl = 0; tl < kSpanBytes / NPOS; ++tl) {
                const int bit = 2 * NPOS * tl, wi = bit >> 5, sh = bit & 31;
                const uint32_t x = sh == 0 ? W[wi] : __funnelshift_r(W[wi], W[wi + 1], sh);
                const uint32_t e = sIndex[x & 0xFFFFu];
                const int pos = NPOS * tl;
                if (pos < 32) {
                    candLo |= e << pos;
                    if (pos + NPOS > 32) candHi |= e >> (32 - pos);
                } else {
                    candHi |= e << (pos - 32);
                }
            }
            candHi &= 0xFFFFu;
            }

            // ---- compaction: scan the row's candidate counts, scatter into the queue ----------------------
            auto scatter = [&](uint32_t lo, uint32_t hi, uint32_t o) {
                while (lo) { const int b = __ffs(lo) - 1; lo &= lo - 1; sQueue[o++] = static_cast<uint16_t>(spanByte + b); }
                while (hi) { const int b = __ffs(hi) - 1; hi &= hi - 1; sQueue[o++] = static_cast<uint16_t>(spanByte + 32 + b); }
            };
            const uint32_t nCand = __popc(candLo) + __popc(candHi);
            const uint32_t incl = warpInclusiveScan(nCand);
            const uint32_t allTotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
            __syncwarp();
            if (allTotal <= static_cast<uint32_t>(kQueueCap)) {
                scatter(candLo, candHi, incl - nCand);
                __syncwarp();
                for (uint32_t from = 0; from < allTotal; from += 32) resolveBatch(from, min(32u, allTotal - from));
            } else {
                // dense row: 11 lanes at a time (<= 528 candidates per step); lane groups keep genome order
                for (int part = 0; part < 3; ++part) {
                    const bool mine = lane >= 11 * part && lane < 11 * (part + 1);
                    const uint32_t lo = mine ? candLo : 0u, hi = mine ? candHi : 0u;
                    const uint32_t n = __popc(lo) + __popc(hi);
                    const uint32_t in2 = warpInclusiveScan(n);
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, in2, 31);
                    __syncwarp();
                    scatter(lo, hi, in2 - n);
                    __syncwarp();
                    for (uint32_t from = 0; from < total; from += 32) resolveBatch(from, min(32u, total - from));
                }
            }
            if (lane == 0) {
                p.slotMeta[slotIdx] = make_uint4(wMatches, ovBase, rowCan, rowNct);
                if (wMatches) atomicAdd(p.tileTot + tileIdx, wMatches);
                if (rowCan) atomicAdd(p.tileTot + static_cast<size_t>(p.nTiles) + tileIdx, rowCan);
                if (rowNct) atomicAdd(p.tileTot + 2 * static_cast<size_t>(p.nTiles) + tileIdx, rowNct);
            }
        }

        // ---- tile epilogue: flush the shared A/C/G/T counters ---------------------------------------------
        __syncthreads();
        if (COUNT) {
            for (int i = tid; i < kWindowSlots * 4; i += kScanThreads) {
                const uint32_t v = sWin[i];
                if (v) atomicAdd(&p.windows[static_cast<size_t>(t.firstWin + (i >> 2)) * 8 + (i & 3)], v);
            }
        }
        tCur = tNext; tNext = tNext2; rCur = rNext;
        tileCur = tileNext; tileNext = tileNext2; rowsCur = rowsNext; rowsNext = rowsNext2;
    }
}

size_t scanSmemBytes() {
    return 65536 + kFlagTabBytes + 2 * (kTileSmemBytes + kTileHalo) + kScanWarps * kQueueCap * 2 +
           2 * kRegionCache * sizeof(DevRegion) + 2 * kWindowSlots * 4 * 4 + 2 * sizeof(TileMeta) + 2 * 8 + 64;
}

cudaError_t launchScan(const ScanParams &p, int grid, cudaStream_t stream) {
    const size_t smem = scanSmemBytes();
    const bool count = p.cfg.countAcgt != 0;
    void (*k)(const ScanParams) = nullptr;
    switch (p.cfg.npos) {
        case 1: k = count ? scanKernel<1, true> : scanKernel<1, false>; break;
        case 2: k = count ? scanKernel<2, true> : scanKernel<2, false>; break;
        case 3: k = count ? scanKernel<3, true> : scanKernel<3, false>; break;
        default: k = count ? scanKernel<4, true> : scanKernel<4, false>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k<<<grid, kScanThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tls
