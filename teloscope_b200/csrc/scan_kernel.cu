// Fused telomeric motif scan for sm_100a.
//
// One persistent CTA per SM pulls tiles (<= 49136 bases of the staged sequence buffer) from a
// dynamic scheduler.  Per tile:
//   1. the raw ASCII bytes are staged into shared memory with cp.async.bulk (TMA bulk copy,
//      mbarrier completion), double buffered so the next tile loads while this one is scanned;
//   2. every lane packs its 48-base span to 2-bit codes in registers (SWAR multiply-gather) and
//      probes a 64 KiB shared-memory table with the 8-mer at every NPOS-th offset; one byte of
//      the table answers NPOS start positions at once (replaces the reference's Trie walk,
//      include/teloscope.h:15-86, src/teloscope.cpp:437-445 and :546-554);
//   3. candidate positions are compacted per warp (scan + scatter into a shared-memory queue) and
//      resolved 32 at a time: exact ASCII validity, window / region truncation
//      (src/teloscope.cpp:439,:549), orientation + canonical flags, per-window covered-bp and
//      per-item counters;
//   4. A/C/G/T window counts come from popcounts on the packed codes (src/teloscope.cpp:418-435);
//   5. the tile's match records are written in genome order to one compact list: per-warp staging
//      in an L2-resident scratch, then a decoupled look-back over tile totals gives the tile's
//      base offset and the warps copy their records out.
// No tensor cores: this is byte/integer streaming work bound by instruction issue and HBM.
#include <cuda_runtime.h>

#include <cstdint>

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
__device__ __forceinline__ unsigned long long ldAcquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelease(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr unsigned long long kFlagAgg = 1ull << 62;     // tile total published
constexpr unsigned long long kFlagPrefix = 2ull << 62;  // inclusive prefix published
constexpr unsigned long long kValueMask = (1ull << 62) - 1;

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

// Nonzero iff some byte of w is not one of ACGTacgt.  Per byte (b7..b0), after ignoring the case
// bit b5, the four letters are exactly: b7=0, b6=1, b3=0, b0 != b4, (b2 & ~b1) == b4.
// The two relations are evaluated at bit 4 of every byte from left-shifted copies.
__device__ __forceinline__ uint32_t badBytes(uint32_t w) {
    const uint32_t l2 = w << 2, l3 = w << 3, l4 = w << 4;
    const uint32_t rel = (~(l4 ^ w)) | ((l2 & ~l3) ^ w);          // bit 4: any relation violated
    return (rel & 0x10101010u) | ((w ^ 0x40404040u) & 0xC8C8C8C8u);
}

__device__ __forceinline__ bool isBase(uint32_t c) {
    c &= 0xDFu;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

__device__ __forceinline__ uint32_t evenMaskBelow(int n) {   // even bits 0,2,..,2(n-1); n in [0,16]
    return n >= 16 ? 0x55555555u : (((1u << (2 * n)) - 1u) & 0x55555555u);
}

__device__ __forceinline__ uint32_t divWindow(const DevConfig &c, uint32_t rel) {
    const uint32_t t = __umulhi(c.winMagic, rel);
    return (t + ((rel - t) >> c.winSh1)) >> c.winSh2;
}

struct TileCtx {
    uint64_t bufStart;
    uint32_t len;
    uint32_t skew;          // tile byte 0 sits at shared-memory byte `skew`
    uint32_t firstRegion;
    uint32_t nCached;       // regions of this tile in shared memory
    uint32_t firstWin;      // window index of the first slot
};

struct Located {
    bool ok;
    uint32_t cacheIdx;      // index in the shared region cache or 0xFFFFFFFF
    uint32_t item, rel, itemLen, regionEndRel, winBase;
};

__device__ __forceinline__ Located locate(const ScanParams &p, const TileCtx &t, const DevRegion *sRegions,
                                          int q) {
    Located r;
    r.ok = false;
    r.cacheIdx = 0xFFFFFFFFu;
    if (q < 0 || q >= static_cast<int>(t.len)) return r;
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

// Warp-level "add by key" for keys that are non-decreasing across the active lanes: one leader
// per distinct key receives the sums.  f(key, sum0, sum1) runs on the leader lane only.
template <typename F>
__device__ __forceinline__ void warpAddByKey(uint32_t active, uint32_t key, uint32_t v0, uint32_t v1, F f) {
    const int lane = threadIdx.x & 31;
    while (active) {
        const int leader = __ffs(active) - 1;
        const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, key, leader);
        const uint32_t grp = __ballot_sync(0xFFFFFFFFu, key == k0) & active;
        const uint32_t s0 = __reduce_add_sync(0xFFFFFFFFu, (grp >> lane) & 1u ? v0 : 0u);
        const uint32_t s1 = __reduce_add_sync(0xFFFFFFFFu, (grp >> lane) & 1u ? v1 : 0u);
        if (lane == leader) f(k0, s0, s1);
        active &= ~grp;
    }
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
    DevRegion *sRegions = reinterpret_cast<DevRegion *>(sQueueAll + kScanWarps * kQueueCap);
    uint32_t *sRegCnt = reinterpret_cast<uint32_t *>(sRegions + kRegionCache);        // [cache][5]
    uint32_t *sWin = sRegCnt + kRegionCache * 5;                                        // [slots][8]
    uint32_t *sWarpCount = sWin + kWindowSlots * 8;                                     // [warps]
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sWarpCount + kScanWarps);            // [2]
    unsigned long long *sTileBase = reinterpret_cast<unsigned long long *>(sBar + 2);
    uint32_t *sNext = reinterpret_cast<uint32_t *>(sTileBase + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevConfig &cfg = p.cfg;
    uint16_t *sQueue = sQueueAll + warp * kQueueCap;
    DevMatch *warpScratch = p.scratch + (static_cast<size_t>(blockIdx.x) * kScanWarps + warp) * p.scratchPerWarp;

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

    auto issueLoad = [&](uint32_t tileIdx, int buf) {   // one thread
        const DevTile t = p.tiles[tileIdx];
        const uint64_t a = t.bufStart & ~15ull;
        const uint32_t bytes = (static_cast<uint32_t>(t.bufStart - a) + t.len + kTileHalo + 15u) & ~15u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&sBar[buf], bytes);
        uint8_t *dst = sTile + buf * kTileStride;
        for (uint32_t off = 0; off < bytes; off += 16384u)
            bulkLoad(dst + off, p.seq + a + off, min(16384u, bytes - off), &sBar[buf]);
    };

    if (tid == 0) {
        const uint32_t first = atomicAdd(p.tileCounter, 1u);
        *sNext = first;
        if (first < p.nTiles) issueLoad(first, 0);
    }
    __syncthreads();
    uint32_t cur = *sNext;

    for (uint32_t iter = 0; cur < p.nTiles; ++iter) {
        const int buf = iter & 1;
        uint8_t *tile = sTile + buf * kTileStride;
        __syncthreads();                                   // everyone has read *sNext
        if (tid == 0) {
            const uint32_t nxt = atomicAdd(p.tileCounter, 1u);
            *sNext = nxt;
            if (nxt < p.nTiles) issueLoad(nxt, buf ^ 1);
        }
        // ---- tile metadata --------------------------------------------------------------------
        TileCtx t;
        {
            const DevTile dt = p.tiles[cur];
            t.bufStart = dt.bufStart;
            t.len = dt.len;
            t.skew = static_cast<uint32_t>(dt.bufStart & 15ull);
            t.firstRegion = dt.firstRegion;
            const uint64_t tileEnd = dt.bufStart + dt.len;
            uint32_t n = 0;
            while (n < kRegionCache && dt.firstRegion + n < p.nRegions &&
                   p.regions[dt.firstRegion + n].bufStart < tileEnd)
                ++n;
            t.nCached = n;
            t.firstWin = 0;
            if (cfg.mode == 0 && n > 0) {
                const DevRegion r0 = p.regions[dt.firstRegion];
                const uint64_t s = dt.bufStart > r0.bufStart ? dt.bufStart : r0.bufStart;
                t.firstWin = r0.winBase + divWindow(cfg, r0.relStart + static_cast<uint32_t>(s - r0.bufStart));
            }
        }
        if (tid < static_cast<int>(t.nCached)) sRegions[tid] = p.regions[t.firstRegion + tid];
        for (int i = tid; i < kRegionCache * 5; i += kScanThreads) sRegCnt[i] = 0;
        for (int i = tid; i < kWindowSlots * 8; i += kScanThreads) sWin[i] = 0;
        __syncthreads();
        const uint32_t nextTile = *sNext;
        mbarWait(&sBar[buf], (iter >> 1) & 1u);

        // ---- per-warp state -----------------------------------------------------------------
        uint32_t qCount = 0;        // candidates pending in this warp's queue
        uint32_t wMatches = 0;      // match records this warp has staged for the tile

        auto addWindow = [&](uint32_t win, int field, uint32_t v) {
            if (v == 0) return;
            const uint32_t slot = win - t.firstWin;
            if (slot < kWindowSlots) atomicAdd(&sWin[slot * 8 + field], v);
            else atomicAdd(&p.windows[static_cast<size_t>(win) * 8 + field], v);
        };

        // ---- stage 2: resolve up to 32 queued candidates (entries [from, from+count)) -----------
        auto resolveBatch = [&](uint32_t from, uint32_t count) {
            const bool active = lane < static_cast<int>(count);
            int sb = 0;
            Located loc;
            loc.ok = false;
            if (active) {
                sb = sQueue[from + lane];
                loc = locate(p, t, sRegions, sb - static_cast<int>(t.skew));
            }
            uint32_t hitMask = 0, flagBits = 0;   // per class: hit bit, 3 flag bits
            uint32_t winIdx = 0;
            bool terminal = false;
            if (loc.ok) {
                // how far a match starting here may extend
                uint32_t limit = loc.regionEndRel;
                if (cfg.mode == 0) {
                    winIdx = divWindow(cfg, loc.rel);
                    const uint64_t wEnd = static_cast<uint64_t>(winIdx + 1) * cfg.windowSize;
                    if (wEnd < limit) limit = static_cast<uint32_t>(wEnd);
                    // terminal status for non-canonical matches (src/teloscope.cpp:403-406,451-459)
                    const uint32_t T = cfg.terminalLimit, n = loc.itemLen;
                    const uint64_t wStart = static_cast<uint64_t>(winIdx) * cfg.windowSize;
                    const uint64_t wEndC = wEnd < n ? wEnd : n;
                    const uint32_t termEnd = n > T ? n - T : 0;
                    if (wEndC <= T || wStart >= termEnd) terminal = true;
                    else if (wStart > T && wEndC < termEnd) terminal = false;
                    else terminal = loc.rel <= T || loc.rel >= termEnd;
                }
                const uint32_t room = limit - loc.rel;          // >= 1
                if (room >= cfg.minLen) {
                    // read 8 bytes at sb (unaligned) as two words
                    const uint32_t a = static_cast<uint32_t>(sb) & ~3u, sh = (static_cast<uint32_t>(sb) & 3u) * 8u;
                    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(tile + a);
                    const uint32_t w1 = *reinterpret_cast<const uint32_t *>(tile + a + 4);
                    const uint32_t w2 = *reinterpret_cast<const uint32_t *>(tile + a + 8);
                    const uint32_t x0 = __funnelshift_r(w0, w1, sh), x1 = __funnelshift_r(w1, w2, sh);
                    uint64_t key = pack4(x0) | (pack4(x1) << 8);
                    // number of leading valid bases among the first 8
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
                                const uint64_t mk = p.sortedKeys[mid];
                                if (mk < k) lo = mid + 1; else hi = mid;
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
            // ---- ordered emission into the warp's staging area ------------------------------------
            const uint32_t nHit = __popc(hitMask);
            uint32_t incl = nHit;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (total == 0) return;
            uint32_t nFwd = 0, nCan = 0, nNct = 0, canLen = 0, nonCanLen = 0, fwdLen = 0, revLen = 0;
            {
                uint32_t o = wMatches + incl - nHit;
                uint32_t hm = hitMask;
                while (hm) {
                    const int ci = __ffs(hm) - 1;
                    hm &= hm - 1;
                    const uint32_t fl = (flagBits >> (2 * ci)) & 3u;      // bit0 fwd, bit1 canonical
                    const uint32_t len = cfg.cls[ci].len;
                    const bool fwd = fl & 1u, can = fl & 2u;
                    DevMatch m;
                    m.pos = loc.rel;
                    m.len = static_cast<uint16_t>(len);
                    m.flags = static_cast<uint8_t>((fwd ? kMatchFwd : 0) | (can ? kMatchCan : 0) |
                                                   (terminal ? kMatchTerm : 0));
                    m.pad = 0;
                    if (o < p.scratchPerWarp) warpScratch[o] = m;
                    ++o;
                    nFwd += fwd; nCan += can; nNct += (!can && terminal);
                    if (can) canLen += len; else nonCanLen += len;
                    if (fwd) fwdLen += len; else revLen += len;
                }
            }
            wMatches += total;
            const uint32_t have = __ballot_sync(0xFFFFFFFFu, nHit != 0);
            // per-item counters (aggregated per distinct region, then one shared atomic each)
            {
                const uint32_t key = loc.cacheIdx != 0xFFFFFFFFu ? loc.cacheIdx : (0x80000000u | loc.item);
                warpAddByKey(have, key, nHit | (nFwd << 10) | (nCan << 20), nNct,
                             [&](uint32_t k, uint32_t s0, uint32_t s1) {
                                 const uint32_t all = s0 & 1023u, fw = (s0 >> 10) & 1023u, cn = s0 >> 20;
                                 if (!(k & 0x80000000u)) {
                                     uint32_t *c = sRegCnt + k * 5;
                                     atomicAdd(c + 0, all); atomicAdd(c + 1, fw); atomicAdd(c + 2, all - fw);
                                     atomicAdd(c + 3, cn); atomicAdd(c + 4, s1);
                                 } else {
                                     DevItemCounts *c = p.itemCounts + (k & 0x7FFFFFFFu);
                                     atomicAdd(&c->all, all); atomicAdd(&c->fwd, fw); atomicAdd(&c->rev, all - fw);
                                     atomicAdd(&c->can, cn); atomicAdd(&c->nonCanTerm, s1);
                                 }
                             });
            }
            if (cfg.mode == 0) {   // per-window covered bp (src/teloscope.cpp:486-507)
                const uint32_t win = loc.winBase + winIdx;
                warpAddByKey(have, win, canLen | (nonCanLen << 16), fwdLen | (revLen << 16),
                             [&](uint32_t k, uint32_t s0, uint32_t s1) {
                                 addWindow(k, 4, s0 & 0xFFFFu); addWindow(k, 5, s0 >> 16);
                                 addWindow(k, 6, s1 & 0xFFFFu); addWindow(k, 7, s1 >> 16);
                             });
            }
        };

        auto drain = [&](bool all) {
            uint32_t from = 0;
            while (qCount - from >= 32u) { resolveBatch(from, 32); from += 32; }
            if (all && qCount > from) { resolveBatch(from, qCount - from); from = qCount; }
            const uint32_t rem = qCount - from;
            if (from > 0 && rem > 0) {
                uint16_t v = 0;
                if (lane < static_cast<int>(rem)) v = sQueue[from + lane];
                __syncwarp();
                if (lane < static_cast<int>(rem)) sQueue[lane] = v;
            }
            __syncwarp();
            qCount = rem;
        };

        // push the candidates (bit i = span position i) of every lane, lane order = genome order
        auto push = [&](uint32_t lo, uint32_t hi, uint32_t spanByte) {
            const uint32_t n = __popc(lo) + __popc(hi);
            uint32_t incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint32_t o = qCount + incl - n;
            while (lo) { const int b = __ffs(lo) - 1; lo &= lo - 1; sQueue[o++] = static_cast<uint16_t>(spanByte + b); }
            while (hi) { const int b = __ffs(hi) - 1; hi &= hi - 1; sQueue[o++] = static_cast<uint16_t>(spanByte + 32 + b); }
            __syncwarp();
            qCount += total;
        };

        // ---- main loop over this warp's rows ------------------------------------------------------
        const uint32_t usedBytes = t.skew + t.len;
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
            const uint32_t row = warp * kRowsPerWarp + rr;
            if (row * kRowBytes >= usedBytes) break;
            const uint32_t spanByte = row * kRowBytes + lane * kSpanBytes;
            const uint4 *src = reinterpret_cast<const uint4 *>(tile + spanByte);
            const uint4 c0 = src[0], c1 = src[1], c2 = src[2], c3 = src[3];
            uint32_t W[5];
            W[0] = pack16(c0); W[1] = pack16(c1); W[2] = pack16(c2); W[3] = pack16(c3); W[4] = 0;

            if (COUNT) {
                // ---- A/C/G/T counts of this span, split at a window / region boundary -----------
                const int q0 = static_cast<int>(spanByte) - static_cast<int>(t.skew);
                int lo = q0 < 0 ? -q0 : 0;                                   // first span position inside the tile
                int hi = static_cast<int>(t.len) - q0; hi = hi > kSpanBytes ? kSpanBytes : hi;   // one past last
                uint32_t winA = 0xFFFFFFFFu, winB = 0xFFFFFFFFu;
                uint32_t acA = 0, gtA = 0, acB = 0, gtB = 0;                 // packed A|C<<16, G|T<<16
                if (lo < hi) {
                    const Located a = locate(p, t, sRegions, q0 + lo);
                    const Located z = locate(p, t, sRegions, q0 + hi - 1);
                    const bool simple = a.ok && z.ok && a.item == z.item &&
                                        z.rel - a.rel == static_cast<uint32_t>(hi - 1 - lo);
                    const uint32_t bad = badBytes(c0.x) | badBytes(c0.y) | badBytes(c0.z) | badBytes(c0.w) |
                                         badBytes(c1.x) | badBytes(c1.y) | badBytes(c1.z) | badBytes(c1.w) |
                                         badBytes(c2.x) | badBytes(c2.y) | badBytes(c2.z) | badBytes(c2.w);
                    const uint32_t wa = simple ? divWindow(cfg, a.rel) : 0, wz = simple ? divWindow(cfg, z.rel) : 0;
                    if (simple && bad == 0 && wz - wa <= 1u) {
                        // fast path: all 48 bytes are bases, one item, at most one window boundary
                        int split = kSpanBytes;                                  // span position of the boundary
                        if (wz != wa) split = lo + static_cast<int>(static_cast<uint64_t>(wz) * cfg.windowSize - a.rel);
                        winA = a.winBase + wa;
                        if (wz != wa) winB = a.winBase + wz;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const uint32_t w = W[j], h = w >> 1;
                            const int l = lo - 16 * j, r = hi - 16 * j, s = split - 16 * j;
                            const uint32_t inc = evenMaskBelow(r < 0 ? 0 : (r > 16 ? 16 : r)) &
                                                 ~evenMaskBelow(l < 0 ? 0 : (l > 16 ? 16 : l));
                            const uint32_t mA = inc & evenMaskBelow(s < 0 ? 0 : (s > 16 ? 16 : s));
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
                        // slow path: region edge, several windows, or a non-ACGT byte in the span
                        for (int i = lo; i < hi; ++i) {
                            const uint32_t c = tile[spanByte + i];
                            if (!isBase(c)) continue;
                            const Located m = locate(p, t, sRegions, q0 + i);
                            if (!m.ok) continue;
                            const uint32_t code = (c >> 1) & 3u;                // A0 C1 T2 G3
                            const int field = code == 0 ? 0 : (code == 1 ? 1 : (code == 3 ? 2 : 3));
                            addWindow(m.winBase + divWindow(cfg, m.rel), field, 1);
                        }
                    }
                }
                // reduce per window across the warp (window ids are non-decreasing in lane order)
                const uint32_t hasA = __ballot_sync(0xFFFFFFFFu, winA != 0xFFFFFFFFu);
                warpAddByKey(hasA, winA, acA, gtA, [&](uint32_t k, uint32_t s0, uint32_t s1) {
                    addWindow(k, 0, s0 & 0xFFFFu); addWindow(k, 1, s0 >> 16);
                    addWindow(k, 2, s1 & 0xFFFFu); addWindow(k, 3, s1 >> 16);
                });
                const uint32_t hasB = __ballot_sync(0xFFFFFFFFu, winB != 0xFFFFFFFFu);
                warpAddByKey(hasB, winB, acB, gtB, [&](uint32_t k, uint32_t s0, uint32_t s1) {
                    addWindow(k, 0, s0 & 0xFFFFu); addWindow(k, 1, s0 >> 16);
                    addWindow(k, 2, s1 & 0xFFFFu); addWindow(k, 3, s1 >> 16);
                });
            }

            // ---- stage 1: one table byte per NPOS start positions ---------------------------------
            uint32_t candLo = 0, candHi = 0;
#pragma unroll
            for (int tl = 0; tl < kSpanBytes / NPOS; ++tl) {
                constexpr int dummy = 0; (void)dummy;
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

            const uint32_t rowTotal = __reduce_add_sync(0xFFFFFFFFu, __popc(candLo) + __popc(candHi));
            if (qCount + rowTotal <= static_cast<uint32_t>(kQueueCap)) {
                push(candLo, candHi, spanByte);
            } else {
                // dense row: one 16-position chunk at a time (<= 512 candidates each)
                drain(false);
                push(candLo & 0xFFFFu, 0, spanByte);
                drain(false);
                push(candLo & 0xFFFF0000u, 0, spanByte);
                drain(false);
                push(0, candHi, spanByte);
            }
            drain(false);
        }
        drain(true);

        // ---- tile epilogue: ordered output -----------------------------------------------------
        if (lane == 0) sWarpCount[warp] = wMatches;
        __syncthreads();
        if (warp == 0) {
            uint32_t c = lane < kScanWarps ? sWarpCount[lane] : 0;
            uint32_t incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += v;
            }
            const unsigned long long tileTotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (lane < kScanWarps) sWarpCount[lane] = incl - c;          // exclusive offsets
            if (lane == 0) {
                unsigned long long excl = 0;
                if (cur == 0) {
                    stRelease(&p.tileStatus[0], kFlagPrefix | tileTotal);
                } else {
                    stRelease(&p.tileStatus[cur], kFlagAgg | tileTotal);
                    for (int j = static_cast<int>(cur) - 1; j >= 0;) {
                        const unsigned long long s = ldAcquire(&p.tileStatus[j]);
                        if ((s >> 62) == 2) { excl += s & kValueMask; break; }
                        if ((s >> 62) == 1) { excl += s & kValueMask; --j; }
                    }
                    stRelease(&p.tileStatus[cur], kFlagPrefix | (excl + tileTotal));
                }
                *sTileBase = excl;
                if (cur == p.nTiles - 1) *p.totalMatches = excl + tileTotal;
            }
        }
        // flush shared counters while the look-back spins
        for (int i = tid; i < kWindowSlots * 8; i += kScanThreads) {
            const uint32_t v = sWin[i];
            if (v) atomicAdd(&p.windows[static_cast<size_t>(t.firstWin + (i >> 3)) * 8 + (i & 7)], v);
        }
        for (int i = tid; i < static_cast<int>(t.nCached) * 5; i += kScanThreads) {
            const uint32_t v = sRegCnt[i];
            if (v) atomicAdd(reinterpret_cast<uint32_t *>(p.itemCounts + sRegions[i / 5].item) + (i % 5), v);
        }
        __syncthreads();
        {
            const unsigned long long base = *sTileBase + sWarpCount[warp];
            const uint32_t n = wMatches < p.scratchPerWarp ? wMatches : p.scratchPerWarp;
            for (uint32_t i = lane; i < n; i += 32) {
                const unsigned long long o = base + i;
                if (o < p.matchCap) p.matches[o] = warpScratch[i];
            }
        }
        cur = nextTile;
    }
}

// Explicit instantiations used by api.cu
#define TLS_INST(N, C) template __global__ void scanKernel<N, C>(const ScanParams);
TLS_INST(1, false) TLS_INST(1, true) TLS_INST(2, false) TLS_INST(2, true)
TLS_INST(3, false) TLS_INST(3, true) TLS_INST(4, false) TLS_INST(4, true)
#undef TLS_INST

size_t scanSmemBytes() {
    return 65536 + kFlagTabBytes + 2 * (kTileSmemBytes + kTileHalo) + kScanWarps * kQueueCap * 2 +
           kRegionCache * sizeof(DevRegion) + kRegionCache * 5 * 4 + kWindowSlots * 8 * 4 + kScanWarps * 4 +
           2 * 8 + 8 + 16;
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
