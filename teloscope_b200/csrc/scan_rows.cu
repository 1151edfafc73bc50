// Window-row telomeric motif scan for sm_100a: the north-star configuration's hot kernel.
//
// Scope: genome-wide mode (Teloscope::analyzeWindow, src/teloscope.cpp:387-534, step == window),
// ONE pattern length L <= 7 held in a direct table, kRowMinWindow <= window <= kRowMaxLen, contigs of
// at least kRowMinWindows windows.  Everything else goes to scan_fast.cu / scan_kernel.cu.
//
// The unit of work is the reference's own unit: one WINDOW.  By src/teloscope.cpp:439 a match never
// crosses a window end and windows tile the contig from its start (:604-637), so a window is an
// independent row: no halo, no boundary inside a row, its 8 counters belong to exactly one warp.
//
//   * warps are autonomous: every warp owns a double-buffered 1 KiB staging area and two mbarriers,
//     stages its next row with cp.async.bulk (TMA bulk copy) while it works on the current one, and
//     never meets another warp at a barrier.  Work = units of 8 consecutive windows, unit u goes to
//     warp u mod (warps in the grid);
//   * a lane owns 32 bases: 2 x LDS.128 of the 16-byte-aligned staging area, SWAR pack to 2-bit codes
//     (one multiply per 4 bases) fused with the exact "all bytes are ACGTacgt" test; the row's
//     alignment skew is removed on the packed words (one funnel shift per word);
//   * 11 probes of the 64 KiB 8-mer table (one byte answers NPOS start offsets, table on a 64 KiB
//     boundary of the shared window so that masking + addressing is one LOP3) give the 32-bit bitmap
//     of match starts; starts whose match would cross the window end are masked (:439);
//   * A/C/G/T: 3 popcounts on bit planes + 2 REDUX per window (:418-435);
//   * matches: warp prefix sum, scatter into a per-warp queue, then 32 at a time: orientation /
//     canonical flags from the 4^L table, per-lane class counters (covered bp = count * L,
//     :470-510), one coalesced 8-byte store into the row's slot of the match arena;
//   * the 8 window counters leave as one 32-byte store; per-tile / per-item totals as a handful of
//     atomics per unit.
//   * a row that holds a byte outside ACGTacgt takes a slow exact path in the same kernel (per-byte
//     validity bitmap: dead starts, bases not counted), so nothing is handed to another kernel.
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
__device__ __forceinline__ void mbarInit(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulkLoad(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint32_t ldsU8(uint32_t addr) {      // the tables never change during the kernel
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t shlClamp(uint32_t v, uint32_t n) {   // v << n, 0 for n >= 32 (PTX semantics)
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
__device__ __forceinline__ uint32_t maskBelow(int n) {   // bits [0, n), n in [0, 32]
    return shlClamp(1u, static_cast<uint32_t>(n)) - 1u;
}
__device__ __forceinline__ int clamp32(int v) { return min(max(v, 0), 32); }

// Exact validity of the bytes fed through add(): see scan_fast.cu (same formula, exhaustively checked
// in tests/test_host_cpu.py::test_validity_formula).
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
__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += u;
    }
    return v;
}

// ---- shared memory layout (offsets inside the dynamic block, which starts kWindowBase bytes into the
// shared window on sm_100; checked at run time) ----------------------------------------------------------
constexpr uint32_t kWindowBase = 0x400;
constexpr uint32_t kPkWords = 66;                                         // 2-bit copy of a row: 64 words + pad
constexpr uint32_t kPkOff = 0;
constexpr uint32_t kQueueOff = kPkOff + kRowWarps * kPkWords * 4;         // 8448
constexpr uint32_t kFlagOff = kQueueOff + kRowWarps * kRowQueueCap * 2;   // + 16384
constexpr uint32_t kBarOff = kFlagOff + kFastFlagBytes;                   // + 16384
constexpr uint32_t kIdxOff = 0x10000u - kWindowBase;
static_assert(kBarOff + kRowWarps * 2 * 8 <= kIdxOff, "front region must end before the 8-mer table");
constexpr uint32_t kStageOff = kIdxOff + 65536u;
constexpr uint32_t kRowSmemTotal = kStageOff + kRowWarps * 2 * kRowStage;
static_assert(kRowSmemTotal <= 232448u, "shared memory budget of one SM");

// Slow exact path of a row that holds a byte outside ACGTacgt (IUPAC letters, '\r', ...): per-lane
// bitmap of such bases in row coordinates and the A/C/G/T counts of the lane's valid bases.  The
// reference's Trie stops at such a byte (include/teloscope.h:27-35,63-66) and analyzeWindow does not
// count it (src/teloscope.cpp:418-435).
__device__ __noinline__ uint32_t slowRow(const uint8_t *row, int nValid, uint32_t *ac, uint32_t *gt) {
    uint32_t inv = 0, a = 0, c = 0, g = 0, t = 0;
    for (int i = 0; i < nValid; ++i) {
        const uint32_t ch = row[i] & 0xDFu;
        if (ch == 'A') ++a;
        else if (ch == 'C') ++c;
        else if (ch == 'G') ++g;
        else if (ch == 'T') ++t;
        else inv |= 1u << i;
    }
    *ac = a | (c << 16);
    *gt = g | (t << 16);
    return inv;
}

}  // namespace

template <int NPOS, bool COUNT>
__global__ void __launch_bounds__(kRowThreads, 1) rowScanKernel(const RowParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevConfig &cfg = p.cfg;

    if ((smemAddr(smem) & 0xFFFFu) != kWindowBase) {       // the table must sit on a 64 KiB boundary of the window
        if (tid == 0 && blockIdx.x == 0) atomicExch(p.status, 1u);
        return;
    }
    const uint32_t idxBase = smemAddr(smem) + kIdxOff;          // low 16 bits are zero
    const uint32_t flagBase = smemAddr(smem) + kFlagOff;
    uint32_t *sPk = reinterpret_cast<uint32_t *>(smem + kPkOff) + warp * kPkWords;
    uint16_t *sQueue = reinterpret_cast<uint16_t *>(smem + kQueueOff) + warp * kRowQueueCap;
    uint8_t *sStage = smem + kStageOff + warp * 2 * kRowStage;
    const uint32_t stageAddr = smemAddr(sStage);
    const uint32_t barAddr = smemAddr(smem + kBarOff) + warp * 16;

    // ---- one-time setup ---------------------------------------------------------------------------
    for (int i = tid; i < 65536 / 16; i += kRowThreads)
        reinterpret_cast<uint4 *>(smem + kIdxOff)[i] = reinterpret_cast<const uint4 *>(p.index8)[i];
    for (uint32_t i = tid; i < p.flagTabBytes / 16; i += kRowThreads)
        reinterpret_cast<uint4 *>(smem + kFlagOff)[i] = reinterpret_cast<const uint4 *>(p.flagTab)[i];
    for (int i = lane; i < 2 * kRowStage / 16; i += 32)        // bytes a short row leaves untouched must be harmless
        reinterpret_cast<uint4 *>(sStage)[i] = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
    if (lane == 0) {
        mbarInit(barAddr, 1);
        mbarInit(barAddr + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t L = p.patLen, W = cfg.windowSize, T = cfg.terminalLimit;
    const uint32_t keyMask = (1u << (2 * L)) - 1u;
    const uint32_t totalWarps = gridDim.x * kRowWarps;
    const uint32_t nUnits = p.nTiles * (kTileRows / kRowsPerUnit);
    const uint8_t *seqBase = p.seq;

    // stage row `r` of tile `d` into buffer `buf` (lane 0 only)
    auto issueRow = [&](const DevRowTile &d, uint32_t r, int buf) {
        const uint64_t g = d.bufStart + static_cast<uint64_t>(r) * W;             // first base of the row
        const uint32_t rel = d.relBase + r * W;
        const uint32_t rowLen = min(W, d.itemLen - rel);
        const uint64_t addr = reinterpret_cast<uint64_t>(seqBase) + g;
        const uint32_t skew = static_cast<uint32_t>(addr) & 15u;
        const uint32_t bytes = (skew + rowLen + 15u) & ~15u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(barAddr + buf * 8, bytes);
        bulkLoad(stageAddr + buf * kRowStage, reinterpret_cast<const void *>(addr - skew), bytes, barAddr + buf * 8);
    };
    auto loadTile = [&](uint32_t u) {
        DevRowTile d;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tiles + (u / (kTileRows / kRowsPerUnit)));
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        d.bufStart = static_cast<uint64_t>(a.x) | (static_cast<uint64_t>(a.y) << 32);
        d.tile = a.z; d.relBase = a.w; d.item = b.x; d.itemLen = b.y; d.firstWin = b.z; d.nRows = b.w;
        return d;
    };

    uint32_t phase = 0;                                        // bit b: parity to wait for on buffer b
    int buf = 0;
    uint32_t u = blockIdx.x * kRowWarps + warp;
    // skip units that lie past the last row of their tile
    DevRowTile d{};
    uint32_t r0 = 0, r1 = 0;
    auto settle = [&]() {                                      // find the next non-empty unit at or after u
        while (u < nUnits) {
            d = loadTile(u);
            r0 = (u % (kTileRows / kRowsPerUnit)) * kRowsPerUnit;
            r1 = min(r0 + kRowsPerUnit, d.nRows);
            if (r0 < r1) return;
            u += totalWarps;
        }
    };
    settle();
    if (u < nUnits && lane == 0) issueRow(d, r0, 0);

    while (u < nUnits) {
        const DevRowTile cur = d;
        const uint32_t rBeg = r0, rEnd = r1;
        // the unit after this one (descriptor loaded now, used for the look-ahead of the last row)
        u += totalWarps;
        settle();
        uint32_t uAll = 0, uCan = 0, uFwd = 0, uNct = 0;

        for (uint32_t r = rBeg; r < rEnd; ++r) {
            // ---- look-ahead: stage the next row of this warp ----------------------------------------------
            __syncwarp();                                       // every lane is done with the other buffer
            if (lane == 0) {
                if (r + 1 < rEnd) issueRow(cur, r + 1, buf ^ 1);
                else if (u < nUnits) issueRow(d, r0, buf ^ 1);
            }
            const uint32_t rel = cur.relBase + r * W;                      // item position of the window start
            const uint32_t rowLen = min(W, cur.itemLen - rel);
            const uint32_t skew = static_cast<uint32_t>(reinterpret_cast<uint64_t>(seqBase) + cur.bufStart +
                                                        static_cast<uint64_t>(r) * W) & 15u;
            const uint8_t *stage = sStage + buf * kRowStage;
            mbarWait(barAddr + buf * 8, (phase >> buf) & 1u);
            phase ^= 1u << buf;

            // ---- 1. load, pack, validate, remove the skew -----------------------------------------------------
            uint32_t X0, X1, X2;
            bool bad;
            {
                const uint4 *src = reinterpret_cast<const uint4 *>(stage + lane * kRowSpan);
                const uint4 c0 = src[0], c1 = src[1];
                ValidAcc va;
                const uint32_t P0 = pack16v(c0, va), P1 = pack16v(c1, va);
                bad = va.bad() != 0 && static_cast<uint32_t>(lane * kRowSpan) < rowLen + skew;
                uint32_t P2 = __shfl_down_sync(0xFFFFFFFFu, P0, 1), P3 = __shfl_down_sync(0xFFFFFFFFu, P1, 1);
                if (lane == 31) { P2 = 0; P3 = 0; }
                const uint32_t sh = 2u * skew;
                X0 = __funnelshift_r(P0, P1, sh);
                X1 = __funnelshift_r(P1, P2, sh);
                X2 = __funnelshift_r(P2, P3, sh);
            }
            const int nValid = clamp32(static_cast<int>(rowLen) - lane * kRowSpan);       // row bases of this lane
            const bool anyBad = __any_sync(0xFFFFFFFFu, bad);
            uint32_t inv = 0, slowAC = 0, slowGT = 0;
            if (anyBad) inv = slowRow(stage + skew + lane * kRowSpan, nValid, &slowAC, &slowGT);
            *reinterpret_cast<uint2 *>(sPk + 2 * lane) = make_uint2(X0, X1);

            // ---- 2. one table byte per NPOS start positions ---------------------------------------------------
            uint32_t H = 0;
            {
                const uint32_t Xv[4] = {X0, X1, X2, 0u};
                constexpr int kProbes = (kRowSpan + NPOS - 1) / NPOS;
#pragma unroll
                for (int tl = 0; tl < kProbes; ++tl) {
                    const int bit = 2 * NPOS * tl, wi = bit >> 5, sh = bit & 31;
                    const uint32_t x = sh == 0 ? Xv[wi] : (sh <= 16 ? Xv[wi] >> sh : __funnelshift_r(Xv[wi], Xv[wi + 1], sh));
                    const uint32_t e = ldsU8((x & 0xFFFFu) | idxBase);
                    H = e * (1u << (NPOS * tl)) + H;              // disjoint bit fields: add == or (bits >= 32 drop out)
                }
            }
            // ---- 3. matches never cross the window end (src/teloscope.cpp:439) --------------------------------
            H &= maskBelow(clamp32(static_cast<int>(rowLen) - static_cast<int>(L) + 1 - lane * kRowSpan));
            if (anyBad) {
                uint32_t invN = __shfl_down_sync(0xFFFFFFFFu, inv, 1);
                if (lane == 31) invN = 0;
                uint32_t dead = inv;
                for (uint32_t j = 1; j < L; ++j) dead |= __funnelshift_r(inv, invN, j);
                H &= ~dead;
            }

            // ---- 4. A/C/G/T of the window ---------------------------------------------------------------------
            uint32_t sumAC = 0, sumGT = 0;
            if (COUNT) {
                uint32_t totAC, totGT;
                const uint32_t Y0 = X0 & (shlClamp(1u, 2u * static_cast<uint32_t>(min(nValid, 16))) - 1u);
                const uint32_t Y1 = X1 & (shlClamp(1u, 2u * static_cast<uint32_t>(max(nValid - 16, 0))) - 1u);
                const uint32_t e = (Y0 & 0x55555555u) | ((Y1 << 1) & 0xAAAAAAAAu);    // code bit 0: C or G
                const uint32_t o = ((Y0 >> 1) & 0x55555555u) | (Y1 & 0xAAAAAAAAu);    // code bit 1: G or T
                const uint32_t nE = __popc(e), nO = __popc(o), nG = __popc(e & o);
                totAC = (static_cast<uint32_t>(nValid) - nE - nO + nG) | ((nE - nG) << 16);
                totGT = nG | ((nO - nG) << 16);
                if (anyBad) { totAC = slowAC; totGT = slowGT; }
                sumAC = __reduce_add_sync(0xFFFFFFFFu, totAC);
                sumGT = __reduce_add_sync(0xFFFFFFFFu, totGT);
            }

            // ---- 5. compaction and emission -------------------------------------------------------------------
            const size_t slotIdx = static_cast<size_t>(cur.tile) * kTileRows + r;
            DevMatch *slot = p.slots + slotIdx * p.slotCap;
            // "terminal" status (src/teloscope.cpp:403-406,451-459) is a property of the window, except for
            // the window that straddles T or n - T
            const uint32_t termEnd = cur.itemLen > T ? cur.itemLen - T : 0u;
            const uint32_t wEnd = rel + rowLen;
            const uint32_t termMode = (wEnd <= T || rel >= termEnd) ? 1u : ((rel > T && wEnd < termEnd) ? 0u : 2u);
            uint32_t wMatches = 0, ovBase = kNoOverflow, rowNct = 0;
            uint32_t cls = 0;                                   // per lane: 4 x 8-bit counters by class (bit0 forward, bit1 canonical)
            auto emitRange = [&](uint32_t total) {              // queue[0, total) holds row positions, ascending
                const bool roomy = wMatches + total <= p.slotCap;           // warp-uniform
                if (!roomy && ovBase == kNoOverflow) {
                    if (lane == 0) ovBase = atomicAdd(p.ovCursor, p.ovChunk);
                    ovBase = __shfl_sync(0xFFFFFFFFu, ovBase, 0);
                }
                uint2 *out = reinterpret_cast<uint2 *>(slot) + wMatches;
                for (uint32_t from = 0; from < total; from += 32) {
                    const uint32_t idx = from + lane;
                    bool nct = false;
                    if (idx < total) {
                        const uint32_t qq = sQueue[idx];
                        const uint32_t wi = qq >> 4, sh = (qq & 15u) * 2u;
                        const uint32_t key = __funnelshift_r(sPk[wi], sPk[wi + 1], sh) & keyMask;
                        const uint32_t f = ldsU8(flagBase + key);        // bit0 hit, bit1 forward, bit2 canonical
                        const uint32_t pos = rel + qq;
                        uint32_t fl = (f >> 1) & 3u;
                        cls += 1u << (8u * fl);
                        if (termMode == 1u || (termMode == 2u && (pos <= T || pos >= termEnd))) {
                            fl |= kMatchTerm;
                            nct = !(f & 4u);
                        }
                        // DevMatch {pos, len:16, flags:8, region:8}
                        const uint2 rec = make_uint2(pos, L | (fl << 16) | (static_cast<uint32_t>(kRegionAccounted) << 24));
                        if (roomy) out[idx] = rec;
                        else {
                            const uint32_t o = wMatches + idx;
                            if (o < p.slotCap) reinterpret_cast<uint2 *>(slot)[o] = rec;
                            else if (static_cast<uint64_t>(ovBase) + (o - p.slotCap) < p.ovCap)
                                reinterpret_cast<uint2 *>(p.ovArena)[ovBase + (o - p.slotCap)] = rec;
                        }
                    }
                    if (termMode) rowNct += __popc(__ballot_sync(0xFFFFFFFFu, nct));
                }
                wMatches += total;
            };
            auto scatter = [&](uint32_t bits, uint32_t o) {
                const uint32_t base = lane * kRowSpan;
                while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; sQueue[o++] = static_cast<uint16_t>(base + b); }
            };
            const uint32_t nH = __popc(H);
            const uint32_t incl = warpInclusiveScan(nH, lane);
            const uint32_t allTotal = __shfl_sync(0xFFFFFFFFu, incl, 31);
            __syncwarp();                                       // the 2-bit copy of the row is in place
            if (allTotal <= static_cast<uint32_t>(kRowQueueCap)) {
                scatter(H, incl - nH);
                __syncwarp();
                emitRange(allTotal);
            } else {
                // dense row: 8 lanes (<= 256 starts) at a time, lane groups keep genome order
                for (int part = 0; part < 4; ++part) {
                    const uint32_t bits = (lane >> 3) == part ? H : 0u;
                    const uint32_t n = __popc(bits);
                    const uint32_t in2 = warpInclusiveScan(n, lane);
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, in2, 31);
                    __syncwarp();
                    scatter(bits, in2 - n);
                    __syncwarp();
                    emitRange(total);
                }
            }

            // ---- 6. the window's 8 counters: one 32-byte store -----------------------------------------------
            const uint32_t cEven = __reduce_add_sync(0xFFFFFFFFu, cls & 0x00FF00FFu);            // rev-non | rev-can << 16
            const uint32_t cOdd = __reduce_add_sync(0xFFFFFFFFu, (cls >> 8) & 0x00FF00FFu);      // fwd-non | fwd-can << 16
            const uint32_t nCan = (cEven >> 16) + (cOdd >> 16), nFwd = (cOdd & 0xFFFFu) + (cOdd >> 16);
            if (lane < 8) {
                uint32_t v;
                if (lane < 4) {
                    const uint32_t pk = lane < 2 ? sumAC : sumGT;
                    v = (lane & 1) ? pk >> 16 : pk & 0xFFFFu;
                } else {
                    const uint32_t n = lane == 4 ? nCan : lane == 5 ? wMatches - nCan : lane == 6 ? nFwd : wMatches - nFwd;
                    v = n * L;                                   // canonical, non-canonical, forward, reverse covered bp
                }
                p.windows[static_cast<size_t>(cur.firstWin + r) * 8 + lane] = v;
            }
            if (lane == 0) {
                p.slotCounts[slotIdx] = wMatches;
                p.slotOv[slotIdx] = ovBase;
                p.slotCan[slotIdx] = nCan;
                p.slotNct[slotIdx] = rowNct;
            }
            uAll += wMatches; uCan += nCan; uFwd += nFwd; uNct += rowNct;
            buf ^= 1;
        }
        // ---- unit epilogue: per-tile and per-item totals ----------------------------------------------------
        if (uAll) {
            if (lane < 3) {
                const uint32_t v = lane == 0 ? uAll : lane == 1 ? uCan : uNct;
                if (v) atomicAdd(p.tileTot + static_cast<size_t>(lane) * p.nTilesAll + cur.tile, v);
            } else if (lane < 8) {
                uint32_t *ic = reinterpret_cast<uint32_t *>(p.itemCounts + cur.item);     // all, fwd, rev, can, nonCanTerm
                const uint32_t v = lane == 3 ? uAll : lane == 4 ? uFwd : lane == 5 ? uAll - uFwd : lane == 6 ? uCan : uNct;
                if (v) atomicAdd(ic + (lane - 3), v);
            }
        }
    }
}

size_t rowSmemBytes() { return kRowSmemTotal; }

bool rowConfigOk(const DevConfig &c) {
    return c.mode == 0 && c.nClasses == 1 && c.cls[0].direct != 0 && c.cls[0].len <= 7 && c.cls[0].len + c.npos <= 9 &&
           c.windowSize >= kRowMinWindow && c.windowSize <= static_cast<uint32_t>(kRowMaxLen) && c.step == c.windowSize;
}

cudaError_t launchRowScan(const RowParams &p, int grid, cudaStream_t stream) {
    const size_t smem = rowSmemBytes();
    const bool count = p.cfg.countAcgt != 0;
    void (*k)(const RowParams) = nullptr;
    switch (p.cfg.npos) {
        case 2: k = count ? rowScanKernel<2, true> : rowScanKernel<2, false>; break;
        case 3: k = count ? rowScanKernel<3, true> : rowScanKernel<3, false>; break;
        default: k = count ? rowScanKernel<4, true> : rowScanKernel<4, false>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k<<<grid, kRowThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tls
