// Lean telomeric motif scan for sm_100a: the north-star configuration's hot kernel.
//
// Scope: genome-wide mode (Teloscope::analyzeWindow, src/teloscope.cpp:387-534, step == window),
// ONE pattern length L <= 7 held in a direct table, window >= 512 bases, tiles that lie inside one
// contig.  The host sends every other tile to the general kernel (scan_kernel.cu), and this kernel
// hands back ("redo list") every tile row in which it sees a byte outside ACGTacgt, so that it
// never needs a slow path of its own.
//
// One persistent CTA of 1024 threads per SM; tile k of a CTA = blockIdx.x + k * gridDim.x.
// Per tile (<= 49136 bases, double buffered in shared memory by cp.async.bulk + mbarrier):
//   every warp owns one row of 1536 bases, every lane a span of 48:
//   1. 3 x LDS.128, SWAR pack to 2-bit codes (one multiply per 4 bases) and, fused with it, the
//      exact "all bytes are ACGTacgt" test (two more multiplies per word, see ValidAcc);
//   2. 16 probes of the 64 KiB 8-mer table (one byte answers 3 start offsets; the table sits on a
//      64 KiB boundary of the shared window so that index masking and addressing are one LOP3)
//      -> 48-bit bitmap of match starts; with a valid row these are exactly the matches of the
//      reference's Trie walk (include/teloscope.h:63-66) before truncation;
//   3. spans that straddle a window end, the tile edge or the contig end clear the starts whose
//      match would cross it (src/teloscope.cpp:439) and split their A/C/G/T counts there;
//   4. A/C/G/T per window: popcounts on bit planes of the codes, REDUX over the lanes of a window,
//      one packed shared atomic per window and row (src/teloscope.cpp:418-435);
//   5. the bitmap is compacted into a per-warp queue; 32 matches at a time get their orientation /
//      canonical flags from the 2-bit copy of the row, bump one class counter of their window
//      (covered bp = count * L at flush time, src/teloscope.cpp:470-510) and are stored, in genome
//      order, to the row's slot of the match arena.
// Integer / byte streaming work: no tensor cores; the bound is instruction issue, then HBM.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_common.cuh"
#include "kernels.hpp"

namespace tls {

namespace {

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
// 48-byte tile descriptor, global -> shared without staging registers (LDGSTS)
__device__ __forceinline__ void copyDescAsync(void *dst, const void *src) {
    const uint32_t d = smemAddr(dst);
    const char *g = static_cast<const char *>(src);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(g + 16) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 32), "l"(g + 32) : "memory");
}
__device__ __forceinline__ void asyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t ldsU8(uint32_t addr) {      // the tables never change during the kernel
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Exact validity of the bytes fed through add(): for c in ACGTacgt bit7=0, bit6=1, bit3=0,
// bit4 = (bit2 & ~bit1) ("is T") and bit0 = ~bit4.  With m = c & 6 in {0,2,4,6} the byte 6*m + 0x40
// has bit7=0, bit6=1 and bit4 set only for m == 4, so (y ^ w) must vanish in bits 7,6,4;
// (w << 4) ^ w must have bit4 set (bit0 != bit4); bit3 is OR-accumulated.  (Exhaustively checked
// against the 256 byte values in tests/test_host_cpu.py::test_validity_formula.)
struct ValidAcc {
    uint32_t a1 = 0, a2 = 0xFFFFFFFFu, any = 0;
    __device__ __forceinline__ void add(uint32_t w, uint32_t m) {
        const uint32_t y = m * 6u + 0x40404040u;
        const uint32_t z = w * 16u;
        a1 |= y ^ w;
        a2 &= z ^ w;
        any |= w;
    }
    __device__ __forceinline__ uint32_t bad() const {
        return (a1 & 0xD0D0D0D0u) | (~a2 & 0x10101010u) | (any & 0x08080808u);
    }
};
// 2-bit codes of 4 bases in the top byte: (c >> 1) & 3 maps A,C,T,G (either case) to 0,1,2,3
__device__ __forceinline__ uint32_t packWord(uint32_t w, ValidAcc &v) {
    const uint32_t m = w & 0x06060606u;
    v.add(w, m);
    return m * 0x00820820u;
}
__device__ __forceinline__ uint32_t pack16v(const uint4 &q, ValidAcc &v) {
    const uint32_t a = packWord(q.x, v), b = packWord(q.y, v), c = packWord(q.z, v), d = packWord(q.w, v);
    return __byte_perm(__byte_perm(a, b, 0x0073), __byte_perm(c, d, 0x0073), 0x5410);
}
__device__ __forceinline__ uint32_t evenMaskBelow(int n) {   // even bits 0,2,..,2(n-1); n clamped to [0,16]
    n = n < 0 ? 0 : n;
    return n >= 16 ? 0x55555555u : (((1u << (2 * n)) - 1u) & 0x55555555u);
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

constexpr int kTileStride = kTileSmemBytes + kTileHalo;
constexpr int kPkBytes = kFastWarps * kFastPackedWords * 4;
constexpr int kQueueBytes = kFastWarps * kFastQueueCap * 2;
constexpr int kWinBytes = 2 * kWindowSlots * 8 * 4;
constexpr int kRowWords = kRowBytes / 16;       // packed words per row (96)
// offsets inside the dynamic shared memory block
constexpr uint32_t kWindowBase = 0x400;         // where dynamic shared memory starts in the shared window
constexpr uint32_t kTile0Off = 0;
constexpr uint32_t kPkOff = kTile0Off + kTileStride;
constexpr uint32_t kIdxOff = 0x10000u - kWindowBase;
static_assert(kPkOff + kPkBytes <= kIdxOff, "tile 0 and the 2-bit rows must fit in front of the 8-mer table");
constexpr uint32_t kFlagOff = kIdxOff + 65536u;
constexpr uint32_t kTile1Off = kFlagOff + kFastFlagBytes;
constexpr uint32_t kQueueOff = kTile1Off + kTileStride;
constexpr uint32_t kWinOff = kQueueOff + kQueueBytes;
constexpr uint32_t kMetaOff = kWinOff + kWinBytes;
constexpr uint32_t kMiscOff = kMetaOff + 4 * sizeof(DevFastTile);      // ring of 4 tile descriptors
constexpr uint32_t kFastSmemTotal = kMiscOff + 64;
static_assert(kFastSmemTotal <= 232448u, "shared memory budget of one SM");

}  // namespace

// Shared window slot (8 words): [0] A | C << 16, [1] G | T << 16 (counts of this tile only, < 2^16),
// [2..5] matches by class (bit0 forward, bit1 canonical), [6..7] unused.
template <int NPOS, bool COUNT>
__global__ void __launch_bounds__(kFastThreads, 1) fastScanKernel(const FastParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevConfig &cfg = p.cfg;

    // ---- shared memory layout (compile-time): the 8-mer table must start on a 64 KiB boundary of the
    // shared window so that (index & 0xFFFF) | base addresses it; dynamic shared memory starts
    // kWindowBase bytes into the window on sm_100 (checked below)
    if ((smemAddr(smem) & 0xFFFFu) != kWindowBase) {
        if (tid == 0 && blockIdx.x == 0) atomicExch(p.redoCount + 1, 1u);
        return;
    }
    const uint32_t idxBase = smemAddr(smem) + kIdxOff;          // low 16 bits are zero
    const uint32_t flagBase = smemAddr(smem) + kFlagOff;
    uint32_t *sPk = reinterpret_cast<uint32_t *>(smem + kPkOff) + warp * kFastPackedWords;
    uint16_t *sQueue = reinterpret_cast<uint16_t *>(smem + kQueueOff) + warp * kFastQueueCap;
    uint32_t *sWinAll = reinterpret_cast<uint32_t *>(smem + kWinOff);
    static_assert(kWindowSlots * 8 == kFastThreads, "one thread per window-slot word");
    DevFastTile *sMetaAll = reinterpret_cast<DevFastTile *>(smem + kMetaOff);
    uint64_t *sBar = reinterpret_cast<uint64_t *>(smem + kMiscOff);             // [2]
    uint32_t *sTot = reinterpret_cast<uint32_t *>(smem + kMiscOff + 16);        // [2][4] tile totals: all, canonical, nct

    const uint32_t G = gridDim.x;
    uint32_t tileI = blockIdx.x;
    if (tileI >= p.nTiles) return;

    // ---- one-time setup ---------------------------------------------------------------------------
    for (int i = tid; i < 65536 / 16; i += kFastThreads)
        reinterpret_cast<uint4 *>(smem + kIdxOff)[i] = reinterpret_cast<const uint4 *>(p.index8)[i];
    for (uint32_t i = tid; i < p.flagTabBytes / 16; i += kFastThreads)
        reinterpret_cast<uint4 *>(smem + kFlagOff)[i] = reinterpret_cast<const uint4 *>(p.flagTab)[i];
    if (tid == 0) {
        mbarInit(&sBar[0], 1);
        mbarInit(&sBar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issueLoad = [&](const DevFastTile &t, int buf) {   // one thread
        const uint64_t a = t.bufStart & ~15ull;
        const uint32_t bytes = (static_cast<uint32_t>(t.bufStart - a) + t.len + kTileHalo + 15u) & ~15u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&sBar[buf], bytes);
        uint8_t *dst = smem + (buf ? kTile1Off : kTile0Off);
        for (uint32_t off = 0; off < bytes; off += 16384u)
            bulkLoad(dst + off, p.seq + a + off, min(16384u, bytes - off), &sBar[buf]);
    };

    // tile descriptors travel global -> shared ring (slot = iteration & 3) two iterations ahead
    if (tid == 0) {
        copyDescAsync(&sMetaAll[0], &p.tiles[tileI]);
        if (tileI + G < p.nTiles) copyDescAsync(&sMetaAll[1], &p.tiles[tileI + G]);
        asyncWaitAll();
        issueLoad(sMetaAll[0], 0);
    }
    const uint32_t L = p.patLen;
    const uint32_t keyMask = (1u << (2 * L)) - 1u;

    for (uint32_t iter = 0; tileI < p.nTiles; ++iter, tileI += G) {
        const int buf = iter & 1;
        const uint8_t *tile = smem + (buf ? kTile1Off : kTile0Off);
        uint32_t *sWin = sWinAll + buf * kWindowSlots * 8;
        if (tid == 0) {
            asyncWaitAll();                                                // descriptor of tile k+1 has landed
            if (tileI + G < p.nTiles) issueLoad(sMetaAll[(iter + 1) & 3], buf ^ 1);
            if (tileI + 2 * G < p.nTiles) copyDescAsync(&sMetaAll[(iter + 2) & 3], &p.tiles[tileI + 2 * G]);
            sTot[buf * 4 + 0] = sTot[buf * 4 + 1] = sTot[buf * 4 + 2] = 0;
        }
        sWin[tid] = 0;                                                    // 1024 threads = 128 slots x 8 words
        __syncthreads();
        const DevFastTile &t = sMetaAll[iter & 3];
        const uint32_t tLen = t.len, relBase = t.relBase;
        const uint32_t skew = static_cast<uint32_t>(t.bufStart) & 15u;
        const uint32_t usedBytes = skew + tLen;
        mbarWait(&sBar[buf], (iter >> 1) & 1u);

        const uint32_t rowByte = warp * kRowBytes;
        const size_t slotIdx = static_cast<size_t>(t.tile) * kTileRows + warp;
        uint32_t wMatches = 0, ovBase = kNoOverflow;
        uint32_t rowCan = 0, rowNct = 0;                                 // canonical / terminal non-canonical among them

        if (rowByte < usedBytes) {                                       // warp-uniform: the row holds tile positions
            const uint32_t spanByte = rowByte + lane * kSpanBytes;
            // ---- 1. load, pack, validate ---------------------------------------------------------------
            uint32_t W0, W1, W2, W3;
            bool bad;
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(tile + spanByte);
                const uint4 c0 = src[0], c1 = src[1], c2 = src[2];
                ValidAcc va;
                W0 = pack16v(c0, va); W1 = pack16v(c1, va); W2 = pack16v(c2, va);
                bad = va.bad() != 0 && spanByte < usedBytes + kTileHalo;
            }
            sPk[lane * 3 + 0] = W0; sPk[lane * 3 + 1] = W1; sPk[lane * 3 + 2] = W2;
            {
                uint32_t tail = 0;
                if (lane == 31) {
                    const uint32_t endByte = rowByte + kRowBytes;                 // 8-byte aligned
                    if (endByte < usedBytes + kTileHalo) {
                        const uint2 h = *reinterpret_cast<const uint2 *>(tile + endByte);
                        ValidAcc va;
                        tail = (packWord(h.x, va) >> 24) | ((packWord(h.y, va) >> 24) << 8);
                        if (va.bad()) bad = true;
                    }
                    sPk[kRowWords] = tail;
                }
                const uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, W0, 1);
                W3 = lane == 31 ? tail : nx;
            }
            if (__any_sync(0xFFFFFFFFu, bad)) {
                // a byte outside ACGTacgt: the general kernel redoes this row (slot stays empty here)
                if (lane == 0) {
                    const unsigned int k = atomicAdd(p.redoCount, 1u);
                    if (k < p.redoCap) p.redoList[k] = (t.tile << 5) | static_cast<uint32_t>(warp);
                }
            } else {
                // ---- 2. stage 1: one table byte per NPOS start positions -----------------------------------
                uint32_t candLo = 0, candHi = 0;
                {
                    const uint32_t Wv[5] = {W0, W1, W2, W3, 0u};
#pragma unroll
                    for (int tl = 0; tl < kSpanBytes / NPOS; ++tl) {
                        const int bit = 2 * NPOS * tl, wi = bit >> 5, sh = bit & 31;
                        const uint32_t x = sh == 0 ? Wv[wi] : (sh <= 16 ? Wv[wi] >> sh : __funnelshift_r(Wv[wi], Wv[wi + 1], sh));
                        const uint32_t e = ldsU8((x & 0xFFFFu) | idxBase);
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
                // ---- 3. lane classification: plain spans need no masking -----------------------------------
                const uint32_t regionEndQ = t.regionEndQ;
                const uint32_t candLimitQ = min(tLen, regionEndQ >= L ? regionEndQ - (L - 1) : 0u);
                const int q0 = static_cast<int>(spanByte) - static_cast<int>(skew);
                const uint32_t relA = relBase + static_cast<uint32_t>(q0);
                const uint32_t winA = divWindow(cfg, relA);
                const uint32_t toNext = (winA + 1) * cfg.windowSize - relA;    // exact in 32-bit wrap arithmetic
                const bool plain = q0 >= 0 && q0 + kSpanBytes <= static_cast<int>(candLimitQ) &&
                                   toNext >= static_cast<uint32_t>(kSpanBytes) + L - 1;
                uint32_t totAC = 0, totGT = 0;                                   // A | C << 16, G | T << 16 of the span
                if (COUNT) {
                    // ---- 4. A/C/G/T: bit planes of 32 + 16 bases -----------------------------------------------
                    const uint32_t e01 = (W0 & 0x55555555u) | ((W1 << 1) & 0xAAAAAAAAu);   // code bit 0: C or G
                    const uint32_t o01 = ((W0 >> 1) & 0x55555555u) | (W1 & 0xAAAAAAAAu);   // code bit 1: G or T
                    const uint32_t e2 = W2 & 0x55555555u, o2 = (W2 >> 1) & 0x55555555u;
                    const uint32_t nE = __popc(e01) + __popc(e2), nO = __popc(o01) + __popc(o2);
                    const uint32_t nG = __popc(e01 & o01) + __popc(e2 & o2);
                    totAC = (kSpanBytes - nE - nO + nG) | ((nE - nG) << 16);
                    totGT = nG | ((nO - nG) << 16);
                    const uint32_t ws = winA + t.winDelta;
                    const uint32_t pm = __ballot_sync(0xFFFFFFFFu, plain);
                    if (pm) {
                        const uint32_t gFirst = __shfl_sync(0xFFFFFFFFu, ws, __ffs(pm) - 1);
                        const uint32_t gLast = __shfl_sync(0xFFFFFFFFu, ws, 31 - __clz(pm));
                        for (uint32_t g = gFirst; g <= gLast; ++g) {
                            const bool mine = plain && ws == g;
                            const uint32_t s0 = __reduce_add_sync(0xFFFFFFFFu, mine ? totAC : 0u);
                            const uint32_t s1 = __reduce_add_sync(0xFFFFFFFFu, mine ? totGT : 0u);
                            if (lane < 2) atomicAdd(sWin + g * 8 + lane, lane ? s1 : s0);
                        }
                    }
                }
                if (!plain) {
                    // ---- 3b/4b. fix-up: window end, tile edge or contig end inside (or just after) the span ------
                    auto prefix = [&](int x, uint32_t &ac, uint32_t &gt) {   // A/C/G/T of span positions [0, x)
                        const uint32_t m0 = evenMaskBelow(x), m1 = evenMaskBelow(x - 16), m2 = evenMaskBelow(x - 32);
                        const uint32_t e0 = W0 & m0, e1 = W1 & m1, e2 = W2 & m2;
                        const uint32_t o0 = (W0 >> 1) & m0, o1 = (W1 >> 1) & m1, o2 = (W2 >> 1) & m2;
                        const uint32_t nE = __popc(e0) + __popc(e1) + __popc(e2), nO = __popc(o0) + __popc(o1) + __popc(o2);
                        const uint32_t nG = __popc(e0 & o0) + __popc(e1 & o1) + __popc(e2 & o2);
                        ac = (static_cast<uint32_t>(x) - nE - nO + nG) | ((nE - nG) << 16);
                        gt = nG | ((nO - nG) << 16);
                    };
                    auto below = [](int k, uint32_t &mLo, uint32_t &mHi) {    // 48-bit mask of positions < k (k in [0, 48])
                        mLo = k >= 32 ? 0xFFFFFFFFu : (1u << k) - 1u;
                        mHi = k <= 32 ? 0u : (1u << (k - 32)) - 1u;
                    };
                    if (q0 >= 0 && q0 + kSpanBytes <= static_cast<int>(candLimitQ)) {
                        // common case: the whole span is inside the tile and the contig; the next window
                        // starts at span position pb in [1, 48 + L - 2]
                        const int pb = static_cast<int>(toNext);
                        // matches never cross the window end (src/teloscope.cpp:439): starts in [pb-L+1, pb) are dead
                        const int ds = pb - static_cast<int>(L) + 1 < 0 ? 0 : pb - static_cast<int>(L) + 1;
                        const int de = pb > kSpanBytes ? kSpanBytes : pb;
                        uint32_t aLo, aHi, bLo, bHi;
                        below(de, aLo, aHi); below(ds, bLo, bHi);
                        candLo &= ~(aLo & ~bLo);
                        candHi &= ~(aHi & ~bHi);
                        if (COUNT) {
                            uint32_t *w = sWin + (winA + t.winDelta) * 8;
                            if (pb >= kSpanBytes) {
                                atomicAdd(w + 0, totAC); atomicAdd(w + 1, totGT);
                            } else {
                                uint32_t acSp, gtSp;
                                prefix(pb, acSp, gtSp);
                                atomicAdd(w + 0, acSp); atomicAdd(w + 1, gtSp);
                                atomicAdd(w + 8, totAC - acSp); atomicAdd(w + 9, totGT - gtSp);
                            }
                        }
                    } else {
                        const int lo = q0 < 0 ? -q0 : 0;
                        int hi = static_cast<int>(tLen) - q0; hi = hi > kSpanBytes ? kSpanBytes : hi;
                        int keepHi = static_cast<int>(candLimitQ) - q0; keepHi = keepHi > kSpanBytes ? kSpanBytes : keepHi;
                        uint32_t keepLo = 0, keepHi32 = 0;
                        if (lo < hi) {
                            const uint32_t relLo = relBase + static_cast<uint32_t>(q0 + lo);
                            const uint32_t winLo = divWindow(cfg, relLo);
                            const uint32_t dist = (winLo + 1) * cfg.windowSize - relLo;     // >= 1
                            const int pb = dist < 4096u ? lo + static_cast<int>(dist) : (1 << 20);   // next window's first base
                            uint32_t aLo, aHi, bLo, bHi;
                            if (keepHi > lo) {
                                below(keepHi, aLo, aHi); below(lo, bLo, bHi);
                                keepLo = aLo & ~bLo; keepHi32 = aHi & ~bHi;
                            }
                            const int ds = pb - static_cast<int>(L) + 1 < 0 ? 0 : pb - static_cast<int>(L) + 1;
                            if (ds < kSpanBytes) {
                                const int de = pb > kSpanBytes ? kSpanBytes : pb;
                                below(de, aLo, aHi); below(ds, bLo, bHi);
                                keepLo &= ~(aLo & ~bLo); keepHi32 &= ~(aHi & ~bHi);
                            }
                            if (COUNT) {
                                const int split = pb < hi ? pb : hi;
                                uint32_t acLo = 0, gtLo = 0, acSp, gtSp, acHi, gtHi;
                                if (lo > 0) prefix(lo, acLo, gtLo);
                                if (hi == kSpanBytes) { acHi = totAC; gtHi = totGT; } else prefix(hi, acHi, gtHi);
                                if (split == hi) { acSp = acHi; gtSp = gtHi; } else prefix(split, acSp, gtSp);
                                uint32_t *w = sWin + (winLo + t.winDelta) * 8;
                                atomicAdd(w + 0, acSp - acLo);
                                atomicAdd(w + 1, gtSp - gtLo);
                                if (split < hi) { atomicAdd(w + 8, acHi - acSp); atomicAdd(w + 9, gtHi - gtSp); }
                            }
                        }
                        candLo &= keepLo;
                        candHi &= keepHi32;
                    }
                }

                // ---- 5. compaction and emission -----------------------------------------------------------
                const uint32_t rowRel0 = relBase + rowByte - skew;           // item position of row byte 0 (wraps for q < 0)
                const uint32_t termMixed = t.termMixed, itemLen = t.itemLen, winDelta = t.winDelta;
                DevMatch *slot = p.slots + slotIdx * p.slotCap;
                auto emitRange = [&](uint32_t total) {          // queue[0, total) holds row byte offsets, ascending
                    const bool roomy = wMatches + total <= p.slotCap;           // warp-uniform
                    if (!roomy && ovBase == kNoOverflow) {
                        if (lane == 0) ovBase = takeOverflowChunk(p.ovCursor, p.ovChunk);
                        ovBase = __shfl_sync(0xFFFFFFFFu, ovBase, 0);
                    }
                    uint2 *out = reinterpret_cast<uint2 *>(slot) + wMatches;
                    for (uint32_t from = 0; from < total; from += 32) {
                        const uint32_t idx = from + lane;
                        bool nct = false, can = false;
                        if (idx < total) {
                            const uint32_t qq = sQueue[idx];
                            const uint32_t wi = qq >> 4, sh = (qq & 15u) * 2u;
                            const uint32_t key = __funnelshift_r(sPk[wi], sPk[wi + 1], sh) & keyMask;
                            const uint32_t f = ldsU8(flagBase + key);        // bit0 hit, bit1 forward, bit2 canonical
                            const uint32_t rel = rowRel0 + qq;
                            const uint32_t winIdx = divWindow(cfg, rel);
                            uint32_t fl = (f >> 1) & 3u;
                            can = (f & 4u) != 0;
                            atomicAdd(sWin + (winIdx + winDelta) * 8 + 2 + fl, 1u);
                            if (termMixed && isTerminalMatch(cfg, rel, winIdx, itemLen)) { fl |= kMatchTerm; nct = !(f & 4u); }
                            // DevMatch {pos, len:16, flags:8, region:8}
                            const uint2 rec = make_uint2(rel, L | (fl << 16) | (static_cast<uint32_t>(kRegionAccounted) << 24));
                            if (roomy) out[idx] = rec;
                            else {
                                const uint32_t o = wMatches + idx;
                                if (o < p.slotCap) reinterpret_cast<uint2 *>(slot)[o] = rec;
                                else if (static_cast<uint64_t>(ovBase) + (o - p.slotCap) < p.ovCap)
                                    reinterpret_cast<uint2 *>(p.ovArena)[ovBase + (o - p.slotCap)] = rec;
                            }
                        }
                        rowCan += __popc(__ballot_sync(0xFFFFFFFFu, can));
                        if (termMixed) rowNct += __popc(__ballot_sync(0xFFFFFFFFu, nct));
                    }
                    wMatches += total;
                };
                auto scatter = [&](uint32_t lo, uint32_t hi, uint32_t o) {
                    const uint32_t base = lane * kSpanBytes;
                    while (lo) { const int b = __ffs(lo) - 1; lo &= lo - 1; sQueue[o++] = static_cast<uint16_t>(base + b); }
                    while (hi) { const int b = __ffs(hi) - 1; hi &= hi - 1; sQueue[o++] = static_cast<uint16_t>(base + 32 + b); }
                };
                const uint32_t nCand = __popc(candLo) + __popc(candHi);
                const uint32_t incl = warpInclusiveScan(nCand);
                const uint32_t allTotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
                __syncwarp();
                if (allTotal <= static_cast<uint32_t>(kFastQueueCap)) {
                    scatter(candLo, candHi, incl - nCand);
                    __syncwarp();
                    emitRange(allTotal);
                } else {
                    // dense row: 6 spans (<= 288 starts) at a time, lane groups keep genome order
                    for (int part = 0; part < 6; ++part) {
                        const bool mine = lane >= 6 * part && lane < 6 * (part + 1);
                        const uint32_t lo = mine ? candLo : 0u, hi = mine ? candHi : 0u;
                        const uint32_t n = __popc(lo) + __popc(hi);
                        const uint32_t in2 = warpInclusiveScan(n);
                        const uint32_t total = __shfl_sync(0xFFFFFFFFu, in2, 31);
                        __syncwarp();
                        scatter(lo, hi, in2 - n);
                        __syncwarp();
                        emitRange(total);
                    }
                }
            }
        }
        if (lane == 0) {
            p.slotMeta[slotIdx] = make_uint4(wMatches, ovBase, rowCan, rowNct);
            if (wMatches) atomicAdd(&sTot[buf * 4 + 0], wMatches);
            if (rowCan) atomicAdd(&sTot[buf * 4 + 1], rowCan);
            if (rowNct) atomicAdd(&sTot[buf * 4 + 2], rowNct);
        }

        // ---- tile epilogue: shared window slots -> global counters -----------------------------------------
        __syncthreads();
        {
            const uint32_t slotW = tid >> 3, k = tid & 7;
            const uint32_t *ws = sWin + slotW * 8;
            uint32_t v, n = 0;
            if (k < 4) {
                const uint32_t pk = ws[k >> 1];
                v = (k & 1) ? pk >> 16 : pk & 0xFFFFu;
            } else {
                const uint32_t c0 = ws[2], c1 = ws[3], c2 = ws[4], c3 = ws[5];   // rev-non, fwd-non, rev-can, fwd-can
                n = k == 4 ? c2 + c3 : k == 5 ? c0 + c1 : k == 6 ? c1 + c3 : c0 + c2;   // canonical, non-canonical, forward, reverse
                v = n * L;
            }
            if (v) atomicAdd(&p.windows[static_cast<size_t>(t.firstWin + slotW) * 8 + k], v);
            // per-item totals: sum the match counts of this warp's 4 slots
            n += __shfl_xor_sync(0xFFFFFFFFu, n, 8);
            n += __shfl_xor_sync(0xFFFFFFFFu, n, 16);
            if (lane >= 4 && lane < 8 && n) {
                uint32_t *ic = reinterpret_cast<uint32_t *>(p.itemCounts + t.item);     // all, fwd, rev, can, nonCanTerm
                if (lane == 4) { atomicAdd(ic + 0, n); atomicAdd(ic + 3, n); }
                else if (lane == 5) atomicAdd(ic + 0, n);
                else if (lane == 6) atomicAdd(ic + 1, n);
                else atomicAdd(ic + 2, n);
            }
            if (tid < 3) {
                const uint32_t v = sTot[buf * 4 + tid];
                if (v) atomicAdd(p.tileTot + static_cast<size_t>(tid) * p.nTilesAll + t.tile, v);
                if (tid == 2 && v) atomicAdd(reinterpret_cast<uint32_t *>(p.itemCounts + t.item) + 4, v);
            }
        }
    }
}

size_t fastSmemBytes() { return kFastSmemTotal; }

bool fastConfigOk(const DevConfig &c) {
    return c.mode == 0 && c.nClasses == 1 && c.cls[0].direct != 0 && c.cls[0].len <= 7 && c.cls[0].len + c.npos <= 9 &&
           c.windowSize >= kFastMinWindow && c.step == c.windowSize;
}

cudaError_t launchFastScan(const FastParams &p, int grid, cudaStream_t stream) {
    const size_t smem = fastSmemBytes();
    const bool count = p.cfg.countAcgt != 0;
    void (*k)(const FastParams) = nullptr;
    switch (p.cfg.npos) {
        case 2: k = count ? fastScanKernel<2, true> : fastScanKernel<2, false>; break;
        case 3: k = count ? fastScanKernel<3, true> : fastScanKernel<3, false>; break;
        default: k = count ? fastScanKernel<4, true> : fastScanKernel<4, false>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k<<<grid, kFastThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tls
