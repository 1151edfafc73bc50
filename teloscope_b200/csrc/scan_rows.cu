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
//   * warps are autonomous and never meet at a barrier.  One CTA of 768 threads per SM (24 warps, 71 registers each:
//     fewer, fatter warps beat 32 warps at 64 registers).  Work = units of 16 consecutive windows, unit u goes to
//     warp u mod (warps in the grid);
//   * a lane owns 32 bases of the row: two 16-byte streaming loads (ld.global.nc, no L1 allocation)
//     from the 16-byte aligned span of the row straight into registers.  They are packed at the top of
//     the row's iteration and the same registers are re-armed at once with the NEXT row's loads, which
//     then have the whole iteration to land (software pipeline of depth one row per warp, 24 KiB in
//     flight per SM); no shared-memory staging, no mbarrier;
//   * SWAR pack to 2-bit codes (one multiply per 4 bases) fused with the exact "all bytes are ACGTacgt"
//     test; the row's alignment skew is removed on the packed words (one funnel shift per word);
//   * 11 probes of the 64 KiB 8-mer table (one byte answers NPOS start offsets; the table sits at
//     shared address 0x10000, so the probe is mask + LDS with an immediate) give the 32-bit bitmap of
//     match starts; starts whose match would cross the window end are masked (:439);
//   * A/C/G/T: 3 popcounts on bit planes, 2 REDUX per window on packed partial sums; the four counts
//     are derived from them once per row in the unit epilogue (:418-435);
//   * matches: warp prefix sum of the per-lane counts, then every lane writes its own 8-byte records,
//     in order, behind those of the lanes before it, straight into the row's slot of the match arena
//     (no queue, no shared copy of the row): the 2L-bit key comes out of the lane's registers, the
//     flag byte from the re-coded 4^L table goes into the record as it is and is summed per lane; one
//     REDUX per row turns the sums into the class counts (covered bp = count * L, :470-510).  A bulk
//     (TMA) store of shared-staged records was measured and is slower at this granularity (about 40
//     extra instructions per row for the fence / elect / uniform-register traffic);
//   * a row's results (8 window counters, slot bookkeeping) wait in shared memory until the unit is
//     done; then lane r writes row r's three 16-byte stores, and the per-tile / per-item totals leave
//     as a handful of atomics per unit;
//   * a row that holds a byte outside ACGTacgt takes a slow exact path in the same kernel (per-byte
//     validity bitmap: dead starts, bases not counted), so nothing is handed to another kernel.
// Integer / byte streaming work: no tensor cores; the bound is instruction issue and the shared-memory
// data pipe (random table probes), then HBM.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_common.cuh"
#include "kernels.hpp"

namespace tls {

namespace {

__device__ __forceinline__ uint32_t smemAddr(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// Streaming loads of a lane's 32 bytes (every base is read once, nothing of it belongs in L1), done only
// when a < b; otherwise the registers keep what they hold.  (Plain ld.global.nc and an L2::256B prefetch hint
// were measured: no difference.)
__device__ __forceinline__ void ldStream32If(uint4 &v0, uint4 &v1, uint64_t addr, uint32_t a, uint32_t b) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.lt.u32 p, %9, %10;\n"
        "@p ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%8];\n"
        "@p ld.global.nc.L1::no_allocate.v4.u32 {%4,%5,%6,%7}, [%8+16];\n"
        "}\n"
        : "+r"(v0.x), "+r"(v0.y), "+r"(v0.z), "+r"(v0.w), "+r"(v1.x), "+r"(v1.y), "+r"(v1.z), "+r"(v1.w)
        : "l"(addr), "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t bfind(uint32_t v) {         // position of the highest set bit (FLO)
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t bmskClamp(int n) {   // bits [0, min(n, 32)), n >= 0 (BMSK)
    uint32_t r;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0u), "r"(static_cast<uint32_t>(n)));
    return r;
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

// ---- shared memory layout: ABSOLUTE addresses in the CTA's shared window.  The dynamic block starts
// kWindowBase bytes into it on sm_100 (checked at run time: status 1 otherwise), so every address below
// is a compile-time constant and goes into the instruction as an immediate. -------------------------------
constexpr uint32_t kWindowBase = 0x400;
constexpr uint32_t kWarpBlk = kRowsPerUnit <= 8 ? 256 : (kRowsPerUnit <= 16 ? 512 : 1024);  // per warp: row results of the unit in flight
constexpr uint32_t kRecOff = 0;                            //   8 x 16 bytes of window / slot results + 8 overflow bases
static_assert(kRecOff + kRowsPerUnit * 20 <= kWarpBlk, "per-warp block");
constexpr uint32_t kWarpBase = kWindowBase;                               // 8 KiB of warp blocks
constexpr uint32_t kFlagBase = kWarpBase + kRowWarps * kWarpBlk;          // 16 KiB flag table
constexpr uint32_t kIdxBase = 0x10000u;                                   // 64 KiB 8-mer table
static_assert(kFlagBase + kFastFlagBytes <= kIdxBase, "front region must end before the 8-mer table");
constexpr uint32_t kRowSmemTotal = kIdxBase + 65536u - kWindowBase;
static_assert(kRowSmemTotal <= 232448u, "shared memory budget of one SM");

// the tables never change during the kernel: plain asm (may be scheduled freely), address = index + immediate
__device__ __forceinline__ uint32_t ldsU8Idx(uint32_t key) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(key), "n"(kIdxBase));
    return v;
}
__device__ __forceinline__ uint32_t ldsU8Flag(uint32_t key) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(key), "n"(kFlagBase));
    return v;
}

// Slow exact path of a row that holds a byte outside ACGTacgt (IUPAC letters, '\r', ...): per-lane
// bitmap of such bases in row coordinates and the counts of the lane's valid bases, in the packed form
// the fast path reduces (x: invalid bitmap, y: nE | nO << 16, z: nG | nValid << 16 with nE = C + G,
// nO = G + T).  The reference's Trie stops at such a byte (include/teloscope.h:27-35,63-66) and
// analyzeWindow does not count it (src/teloscope.cpp:418-435).
__device__ __noinline__ uint4 slowRow(const uint8_t *row, int nValid) {
    uint32_t inv = 0, a = 0, c = 0, g = 0, t = 0;
    for (int i = 0; i < nValid; ++i) {
        const uint32_t ch = __ldg(row + i) & 0xDFu;
        if (ch == 'A') ++a;
        else if (ch == 'C') ++c;
        else if (ch == 'G') ++g;
        else if (ch == 'T') ++t;
        else inv |= 1u << i;
    }
    return make_uint4(inv, (c + g) | ((g + t) << 16), g | ((a + c + g + t) << 16), 0u);
}

// A row that is shorter than the window (the last one of a contig) and / or holds a byte outside ACGTacgt:
// the lane's match starts H (computed with the masks of a full window) and its packed A/C/G/T partial sums
// are redone exactly.  Called by the whole warp.
__device__ __noinline__ uint4 fixRow(uint32_t H, uint32_t X0, uint32_t X1, const uint8_t *laneRow, uint32_t rowLen, uint32_t L,
                                     bool anyBad, bool count) {
    const uint32_t lane = threadIdx.x & 31u;
    const int nLeft = static_cast<int>(rowLen) - static_cast<int>(lane * kRowSpan);
    const uint32_t nv = static_cast<uint32_t>(clamp32(nLeft));
    H &= bmskClamp(max(nLeft - static_cast<int>(L) + 1, 0));
    uint32_t t1 = 0, t2 = 0;
    if (count) {
        const uint32_t Y0 = X0 & bmskClamp(max(2 * nLeft, 0)), Y1 = X1 & bmskClamp(max(2 * nLeft - 32, 0));
        const uint32_t e = (Y0 & 0x55555555u) | ((Y1 << 1) & 0xAAAAAAAAu);
        const uint32_t o = ((Y0 >> 1) & 0x55555555u) | (Y1 & 0xAAAAAAAAu);
        t1 = __popc(e) + (__popc(o) << 16);
        t2 = __popc(e & o) + (nv << 16);
    }
    if (anyBad) {
        const uint4 slow = slowRow(laneRow, static_cast<int>(nv));
        uint32_t invN = __shfl_down_sync(0xFFFFFFFFu, slow.x, 1);
        if (lane == 31) invN = 0;
        uint32_t dead = slow.x;
        for (uint32_t j = 1; j < L; ++j) dead |= __funnelshift_r(slow.x, invN, j);
        H &= ~dead;
        t1 = slow.y; t2 = slow.z;
    }
    return make_uint4(H, t1, t2, 0u);
}

}  // namespace

template <int NPOS, bool COUNT>
__global__ void __launch_bounds__(kRowThreads, 1) rowScanKernel(const RowParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (smemAddr(smem) != kWindowBase) {                   // the layout above is in absolute shared addresses
        if (tid == 0 && blockIdx.x == 0) atomicExch(p.status, 1u);
        return;
    }
    const uint32_t wb = kWarpBase + warp * kWarpBlk;       // this warp's block of row results

    // ---- one-time setup: the 8-mer table as it is; the flag table re-coded for the emission loop:
    //      bit0 forward, bit1 canonical (= DevMatch flags), bit7 canonical again (count field, see below) ------
    for (int i = tid; i < 65536 / 16; i += kRowThreads)
        reinterpret_cast<uint4 *>(smem + (kIdxBase - kWindowBase))[i] = reinterpret_cast<const uint4 *>(p.index8)[i];
    for (uint32_t i = tid; i < p.flagTabBytes / 4; i += kRowThreads) {
        const uint32_t f = reinterpret_cast<const uint32_t *>(p.flagTab)[i];   // per byte: bit0 hit, bit1 forward, bit2 canonical
        reinterpret_cast<uint32_t *>(smem + (kFlagBase - kWindowBase))[i] = ((f >> 1) & 0x03030303u) | ((f & 0x04040404u) << 5);
    }
    __syncthreads();

    const uint32_t L = p.patLen, W = p.cfg.windowSize, T = p.cfg.terminalLimit;
    const uint32_t keyMask = (1u << (2 * L)) - 1u;
    const uint32_t recHi = L | (static_cast<uint32_t>(kRegionAccounted) << 24);    // DevMatch {pos, len:16, flags:8, region:8}
    const uint32_t totalWarps = gridDim.x * kRowWarps;
    const uint32_t nUnits = p.nUnits;
    const uint32_t laneOff = lane * kRowSpan;
    const uint64_t seqBase = reinterpret_cast<uint64_t>(p.seq);
    const uint32_t slotBytes = p.slotCap * 8u;
    const bool lane0 = lane == 0;
    const uint32_t laneLt = (1u << lane) - 1u;

    // Per-lane masks of a row of rowLen bases: the lane's valid 2-bit fields in X0 / X1, its valid match
    // starts (a match never crosses the window end, src/teloscope.cpp:439) and its number of bases << 16.
    // A full window has the same masks in every row; the last window of a contig is shorter and applies
    // its own (a subset) on top.
    struct LaneMasks { uint32_t m0, m1, hm, nv; };
    auto laneMasks = [&](uint32_t rowLen) {
        const int nLeft = static_cast<int>(rowLen) - static_cast<int>(laneOff);
        LaneMasks m;
        m.m0 = bmskClamp(max(2 * nLeft, 0));
        m.m1 = bmskClamp(max(2 * nLeft - 32, 0));
        m.hm = bmskClamp(max(nLeft - static_cast<int>(L) + 1, 0));
        m.nv = static_cast<uint32_t>(clamp32(nLeft)) << 16;
        return m;
    };
    const LaneMasks fm = laneMasks(W);

    auto loadUnit = [&](uint32_t u, uint4 &a, uint4 &b) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.units + u);
        a = __ldg(src); b = __ldg(src + 1);
    };
    // The lane's 32 bytes of the 16-byte aligned span of a row whose first base is at rowAddr:
    // ((rowAddr + laneOff) & ~15) since laneOff is a multiple of 16.  Lanes past the row keep what they
    // hold (everything past the row is masked); nothing is read more than 47 bytes past a row's end.
    uint4 c0 = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u), c1 = c0;
    auto loadRow = [&](uint64_t laneRowAddr, uint32_t len) {
        ldStream32If(c0, c1, laneRowAddr & ~static_cast<uint64_t>(15), laneOff, len + 15u);
    };

    uint32_t u = blockIdx.x * kRowWarps + warp;
    if (u >= nUnits) return;
    uint4 da, db;
    loadUnit(u, da, db);
    loadRow(seqBase + (static_cast<uint64_t>(da.x) | (static_cast<uint64_t>(da.y) << 32)) + laneOff, min(W, da.w - da.z));

    for (;;) {
        // laneAddr: address of the lane's first base of the current row (its low 4 bits are the row's skew)
        uint64_t laneAddr = seqBase + (static_cast<uint64_t>(da.x) | (static_cast<uint64_t>(da.y) << 32)) + laneOff;
        uint32_t rel = da.z;                                   // item position of the window start
        const uint32_t itemLen = da.w;
        const uint32_t termEnd = itemLen > T ? itemLen - T : 0u;   // "terminal": pos <= T || pos >= n - T (src/teloscope.cpp:403-406,451-459)
        const uint32_t nRows = db.z;
        const uint32_t uNext = u + totalWarps;
        uint64_t nextAddr = 0;                                 // the lane's first base of the first row of this warp's next unit
        uint32_t nextLim = 0;                                  // its length + 15 (lanes below it load), 0: there is none
        if (uNext < nUnits) {
            const uint4 na = __ldg(reinterpret_cast<const uint4 *>(p.units + uNext));
            nextAddr = seqBase + (static_cast<uint64_t>(na.x) | (static_cast<uint64_t>(na.y) << 32)) + laneOff;
            nextLim = min(W, na.w - na.z) + 15u;
        }
        uint8_t *slotPtr = reinterpret_cast<uint8_t *>(p.slots) + static_cast<size_t>(db.y) * slotBytes;
        uint32_t ovRows = 0;                                   // bit r: row r spilled into an overflow chunk

#pragma unroll 1
        for (uint32_t r = 0; r < nRows; ++r) {
            const uint32_t remaining = itemLen - rel;
            const bool partial = remaining < W;                 // warp-uniform: the last window of a contig
            const uint32_t skew = static_cast<uint32_t>(laneAddr) & 15u;

            // ---- 1. pack + validate the row that has landed, then re-arm the registers with the next row -------
            uint32_t P0, P1, badBits;
            {
                ValidAcc va;
                P0 = pack16v(c0, va);
                P1 = pack16v(c1, va);
                badBits = va.bad();                             // stale lanes past a short row hold bytes that were valid or already took the slow path
            }
            {
                const bool inUnit = r + 1 < nRows;
                const uint64_t na = inUnit ? laneAddr + W : nextAddr;
                const uint32_t nl = inUnit ? min(W, remaining - W) + 15u : nextLim;
                ldStream32If(c0, c1, na & ~static_cast<uint64_t>(15), laneOff, nl);
            }

            // ---- 2. remove the skew: X0 X1 X2 = the lane's 32 bases and the 16 that follow -------------------
            const uint32_t sh = 2u * skew;
            const uint32_t X0 = __funnelshift_r(P0, P1, sh);
            const uint32_t X1 = __funnelshift_r(P1, __shfl_down_sync(0xFFFFFFFFu, P0, 1), sh);
            const uint32_t X2 = __shfl_down_sync(0xFFFFFFFFu, X0, 1);      // lane 31: bases past any row, masked below
            const bool anyBad = __any_sync(0xFFFFFFFFu, badBits != 0);

            // ---- 3. one table byte per NPOS start positions ---------------------------------------------------
            uint32_t H = 0;
            {
                const uint32_t Xv[4] = {X0, X1, X2, 0u};
                constexpr int kProbes = (kRowSpan + NPOS - 1) / NPOS;
#pragma unroll
                for (int tl = 0; tl < kProbes; ++tl) {
                    const int bit = 2 * NPOS * tl, wi = bit >> 5, s = bit & 31;
                    const uint32_t x = s == 0 ? Xv[wi] : (s <= 16 ? Xv[wi] >> s : __funnelshift_r(Xv[wi], Xv[wi + 1], s));
                    const uint32_t e = ldsU8Idx(x & 0xFFFFu);
                    H = e * (1u << (NPOS * tl)) + H;              // disjoint bit fields: add == or (bits >= 32 drop out)
                }
            }
            // ---- 4. matches never cross the window end (src/teloscope.cpp:439) --------------------------------
            H &= fm.hm;

            // ---- 5. A/C/G/T of the window: packed partial sums (nE = C + G, nO = G + T, nG, bases) -------------
            uint32_t t1 = 0, t2 = 0;
            if (COUNT) {
                const uint32_t Y0 = X0 & fm.m0, Y1 = X1 & fm.m1;
                const uint32_t e = (Y0 & 0x55555555u) | ((Y1 << 1) & 0xAAAAAAAAu);    // code bit 0: C or G
                const uint32_t o = ((Y0 >> 1) & 0x55555555u) | (Y1 & 0xAAAAAAAAu);    // code bit 1: G or T
                t1 = __popc(e) + (__popc(o) << 16);
                t2 = __popc(e & o) + fm.nv;
            }
            if (partial || anyBad) {                            // rare, warp-uniform: the exact path redoes H, t1, t2
                const uint4 fx = fixRow(H, X0, X1, reinterpret_cast<const uint8_t *>(laneAddr), min(W, remaining), L, anyBad, COUNT);
                H = fx.x; t1 = fx.y; t2 = fx.z;
            }
            uint32_t sum1 = 0, sum2 = 0;
            if (COUNT) {
                sum1 = __reduce_add_sync(0xFFFFFFFFu, t1);
                sum2 = __reduce_add_sync(0xFFFFFFFFu, t2);
            }

            // ---- 6. emission: every lane writes its own matches, in order, behind those of the lanes before it.
            //      The match's 2L-bit key comes out of the lane's registers, the flag byte g (bit0 forward, bit1
            //      canonical, bit7 canonical) goes into the record as it is and is summed per lane: a lane has at
            //      most 32 matches, so forward + 2 canonical < 128 stays below the canonical count at bit 7. ----------
            const uint32_t rowEnd = rel + min(W, remaining);
            const bool rowHasTerm = rel <= T || rowEnd > termEnd;                               // warp-uniform
            const bool rowAllTerm = rowEnd <= T + 1u || rel >= termEnd;                         // warp-uniform
            const uint32_t nH = __popc(H);
            uint32_t incl, wMatches;
            if (!__any_sync(0xFFFFFFFFu, nH > 7u)) {
                // at most 7 matches per lane (any row but a homopolymer-like one): the prefix from three independent
                // ballots on the bits of the count instead of five dependent shuffles
                const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, nH & 1u), b1 = __ballot_sync(0xFFFFFFFFu, nH & 2u),
                               b2 = __ballot_sync(0xFFFFFFFFu, nH & 4u);
                incl = __popc(b0 & laneLt) + 2u * __popc(b1 & laneLt) + 4u * __popc(b2 & laneLt) + nH;
                wMatches = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
            } else {
                incl = warpInclusiveScan(nH, lane);
                wMatches = __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            const uint32_t relLane = rel + laneOff;
            uint32_t acc = 0, rowNct = 0;
            if (wMatches <= p.slotCap && (rowAllTerm || !rowHasTerm)) {
                // common case: the slot has room and "terminal" is the same for the whole row
                const uint32_t recHiRow = recHi | (rowAllTerm ? static_cast<uint32_t>(kMatchTerm) << 16 : 0u);
                uint2 *out = reinterpret_cast<uint2 *>(slotPtr) + (incl - nH);
                uint32_t bits = H;
                while (bits) {
                    const uint32_t t = bits & (0u - bits);
                    const uint32_t c = bfind(t);
                    bits ^= t;
                    const bool lowHalf = c < 16u;
                    const uint32_t key = __funnelshift_r(lowHalf ? X0 : X1, lowHalf ? X1 : X2, 2u * c) & keyMask;
                    const uint32_t g = ldsU8Flag(key);
                    *out++ = make_uint2(relLane + c, g * 65536u + recHiRow);
                    acc += g;
                }
            } else {
                uint32_t ovBase = kNoOverflow;
                if (wMatches > p.slotCap) {
                    if (lane0) ovBase = atomicAdd(p.ovCursor, p.ovChunk);
                    ovBase = __shfl_sync(0xFFFFFFFFu, ovBase, 0);
                    if (lane0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wb + kRecOff + kRowsPerUnit * 16 + r * 4), "r"(ovBase) : "memory");
                    ovRows |= 1u << r;
                }
                uint32_t o = incl - nH;
                uint32_t bits = H;
                while (bits) {
                    const uint32_t t = bits & (0u - bits);
                    const uint32_t c = bfind(t);
                    bits ^= t;
                    const bool lowHalf = c < 16u;
                    const uint32_t key = __funnelshift_r(lowHalf ? X0 : X1, lowHalf ? X1 : X2, 2u * c) & keyMask;
                    const uint32_t g = ldsU8Flag(key);
                    const uint32_t pos = relLane + c;
                    const bool term = pos <= T || pos >= termEnd;
                    const uint2 rec = make_uint2(pos, g * 65536u + (term ? static_cast<uint32_t>(kMatchTerm) << 16 : 0u) + recHi);
                    if (o < p.slotCap) reinterpret_cast<uint2 *>(slotPtr)[o] = rec;
                    else if (static_cast<uint64_t>(ovBase) + (o - p.slotCap) < p.ovCap)
                        reinterpret_cast<uint2 *>(p.ovArena)[ovBase + (o - p.slotCap)] = rec;
                    ++o;
                    acc += g;
                    if (term && !(g & 2u)) ++rowNct;
                }
                rowNct = __reduce_add_sync(0xFFFFFFFFu, rowNct);
            }
            // forward + 2 canonical | canonical << 16, summed over the lanes
            const uint32_t flagSum = __reduce_add_sync(0xFFFFFFFFu, (acc & 127u) | ((acc >> 7) << 16));

            // ---- 7. the row's results wait in shared memory for the unit epilogue ----------------------------------
            {   // all four values are warp-uniform: every lane stores them to the same address (one wavefront, no election)
                const uint32_t w2 = wMatches | (rowNct << 12) | (rowAllTerm ? 0x80000000u : 0u);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wb + kRecOff + r * 16), "r"(sum1), "r"(sum2), "r"(w2), "r"(flagSum) : "memory");
            }
            laneAddr += W; rel += W;
            slotPtr += slotBytes;
        }
        // ---- unit epilogue: lane r writes row r's window counters and slot bookkeeping (three 16-byte stores);
        //      per-tile and per-item totals as a handful of atomics ------------------------------------------------
        __syncwarp();
        uint32_t k2 = 0, k3 = 0;                               // all | canonical << 16, forward | terminal non-canonical << 16
        if (lane < nRows) {
            uint32_t s1, s2, w2, fs;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(s1), "=r"(s2), "=r"(w2), "=r"(fs) : "r"(wb + kRecOff + lane * 16));
            uint32_t ov = kNoOverflow;
            if ((ovRows >> lane) & 1u) ov = lds32(wb + kRecOff + kRowsPerUnit * 16 + lane * 4);
            const uint32_t all = w2 & 0xFFFu, can = fs >> 16, fwd = (fs & 0xFFFFu) - 2u * can;
            const uint32_t nct = (w2 >> 31) ? all - can : (w2 >> 12) & 0xFFFu;
            const uint32_t nE = s1 & 0xFFFFu, nO = s1 >> 16, nG = s2 & 0xFFFFu, nV = s2 >> 16;
            uint4 *wp = p.windows + 2 * static_cast<size_t>(db.x + lane);
            wp[0] = make_uint4(nV - nE - nO + nG, nE - nG, nG, nO - nG);        // A, C, G, T
            wp[1] = make_uint4(can * L, (all - can) * L, fwd * L, (all - fwd) * L);
            p.slotMeta[db.y + lane] = make_uint4(all, ov, can, nct);
            k2 = all | (can << 16);
            k3 = fwd | (nct << 16);
        }
        k2 = __reduce_add_sync(0xFFFFFFFFu, k2);               // <= 32 rows x 1002 per half: no carry between the halves
        k3 = __reduce_add_sync(0xFFFFFFFFu, k3);
        const uint32_t uAll = k2 & 0xFFFFu, uCan = k2 >> 16, uFwd = k3 & 0xFFFFu, uNct = k3 >> 16;
        if (uAll) {
            const uint32_t tile = db.y / kTileRows;
            if (lane < 3) {
                const uint32_t v = lane == 0 ? uAll : lane == 1 ? uCan : uNct;
                if (v) atomicAdd(p.tileTot + static_cast<size_t>(lane) * p.nTilesAll + tile, v);
            } else if (lane < 8) {
                uint32_t *ic = reinterpret_cast<uint32_t *>(p.itemCounts + db.w);     // all, fwd, rev, can, nonCanTerm
                const uint32_t v = lane == 3 ? uAll : lane == 4 ? uFwd : lane == 5 ? uAll - uFwd : lane == 6 ? uCan : uNct;
                if (v) atomicAdd(ic + (lane - 3), v);
            }
        }
        if (uNext >= nUnits) break;
        u = uNext;
        loadUnit(u, da, db);
    }
}

// Work units of the window-row kernel, one thread per unit: unit j of a contig covers its windows
// [16j, 16j + 16) (kRowsPerUnit), tile j / 2, row slots (j % 2) * 16 ... of that tile.
__global__ void __launch_bounds__(256) rowPlanKernel(const DevRowRegion *regions, uint32_t nRegions, uint32_t W,
                                                     DevRowUnit *units, uint32_t nUnits) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nUnits) return;
    uint32_t lo = 0, hi = nRegions;                               // last region with unit0 <= u
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (regions[mid].unit0 <= u) lo = mid; else hi = mid;
    }
    const DevRowRegion r = regions[lo];
    const uint32_t j = u - r.unit0;
    const uint32_t w0 = j * kRowsPerUnit;
    const uint32_t nWin = (r.len + W - 1) / W;
    DevRowUnit o;
    o.bufStart = r.bufStart + static_cast<uint64_t>(w0) * W;
    o.rel = r.relStart + w0 * W;
    o.itemLen = r.itemLen;
    o.firstWin = r.winBase + w0;
    o.slot0 = (r.tile0 + w0 / kTileRows) * kTileRows + w0 % kTileRows;
    o.nRows = min(static_cast<uint32_t>(kRowsPerUnit), nWin - w0);
    o.item = r.item;
    units[u] = o;
}

cudaError_t launchRowPlan(const DevRowRegion *regions, uint32_t nRegions, uint32_t windowSize, DevRowUnit *units,
                          uint32_t nUnits, cudaStream_t stream) {
    if (nUnits == 0) return cudaSuccess;
    rowPlanKernel<<<(nUnits + 255) / 256, 256, 0, stream>>>(regions, nRegions, windowSize, units, nUnits);
    return cudaGetLastError();
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
