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
//   * 11 probes of the 128 KiB 8-mer table (one u16 entry answers NPOS start offsets: hit / forward / canonical bit
//     fields; the table sits at shared address 0x10000, so the probe is mask + LDS with an immediate);
//     starts whose match would cross the window end are masked (:439);
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
#include "row_common.cuh"

namespace tls {

namespace {

// ---- shared memory layout: ABSOLUTE addresses in the CTA's shared window.  The dynamic block starts
// kWindowBase bytes into it on sm_100 (checked at run time: status 1 otherwise), so every address below
// is a compile-time constant and goes into the instruction as an immediate. -------------------------------
constexpr uint32_t kWindowBase = 0x400;
constexpr uint32_t kWarpBlk = kRowsPerUnit <= 8 ? 512 : (kRowsPerUnit <= 16 ? 1024 : 2048);  // per warp: row results of the unit in flight
constexpr uint32_t kRecOff = 0;                            //   16 bytes of results per row, then per row: overflow base (u32), terminal non-canonical
constexpr uint32_t kOvOff = kRowsPerUnit * 16;             //   count (u32), and the registers of the lane that holds the row's canonical matches (32 bytes)
constexpr uint32_t kNctOff = kRowsPerUnit * 20;
constexpr uint32_t kStashOff = kRowsPerUnit * 32;
static_assert(kRowsPerUnit * 64 <= kWarpBlk, "per-warp block");
constexpr uint32_t kWarpBase = kWindowBase;                               // warp blocks
constexpr uint32_t kFlagBase = kWarpBase + kRowWarps * kWarpBlk;          // 16 KiB flag table
constexpr uint32_t kIdxBase = 0x10000u;                                   // 64 KiB 8-mer table (one byte per key)
static_assert(kFlagBase + kFastFlagBytes <= kIdxBase, "front region must end before the 8-mer table");
constexpr uint32_t kRowSmemTotal = kIdxBase + 65536u - kWindowBase;
static_assert(kRowSmemTotal <= 232448u, "shared memory budget of one SM");

// the tables never change during the kernel: plain asm (may be scheduled freely), address = index + immediate
__device__ __forceinline__ uint32_t ldsU8Idx(uint32_t key) {             // 16-bit 8-mer key
    uint32_t v;
    asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(key), "n"(kIdxBase));
    return v;
}
__device__ __forceinline__ uint32_t ldsU8Flag(uint32_t key) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(key), "n"(kFlagBase));
    return v;
}

// Layout of the per-lane match accumulators.  One byte of the 8-mer table answers NPOS start offsets: bit fields H (some
// pattern matches there) and F (a forward pattern) of NPOS bits each, and one bit C (one of those matches is the canonical
// pattern, either strand).  G consecutive probes share one 32-bit accumulator (probe j of the group sits FW * j bits up), so
// the whole lane (32 start positions) takes kAcc = 3 accumulators, and a probe costs one multiply-add on the FMA pipe: no
// masks, no shifts on the ALU pipe, which is what bounds this kernel.
template <int NPOS>
struct AccLayout {
    static constexpr int FW = 2 * NPOS + 1;
    static constexpr int G = 32 / FW;                                      // 6 (NPOS 2), 4 (NPOS 3)
    static constexpr int kProbes = (kRowSpan + NPOS - 1) / NPOS;           // 16, 11
    static constexpr int kAcc = (kProbes + G - 1) / G;                     // 3
    __host__ __device__ static constexpr uint32_t field(int f) {           // the bits of field f (0 H, 1 F, 2 C) of all G probes
        uint32_t m = 0;
        for (int j = 0; j < G; ++j) m |= (f == 2 ? 1u : (1u << NPOS) - 1u) << (FW * j + f * NPOS);
        return m;
    }
};

}  // namespace

// Two kinds of work units (16 consecutive windows of one contig, see sparseRange() in device_common.cuh):
//   * DENSE units hold a window that touches a terminal zone (the first / last terminalLimit bases of the contig,
//     src/teloscope.cpp:403-406) or the contig's last window.  Every match becomes a record in the row's slot of the
//     arena (the terminal-block walk and the terminal non-canonical list of -m need them all).
//   * SPARSE units (everything else: 98 % of a chromosome) emit only what the reference ever reads again from
//     analyzeWindow's match vectors there: the canonical matches (canonicalMatches, and the seeds of the interstitial
//     blocks).  The other matches only count: window counters, per-item totals.  Per row a 32-bit mask of the lanes
//     without a match start goes out as well; the interstitial-block kernels use it to prove that two canonical
//     matches cannot belong to one run (block_kernels.cu) and rescan the sequence where they cannot.
template <int NPOS, bool COUNT>
__global__ void __launch_bounds__(kRowThreads, 1) rowScanKernel(const RowParams p) {
    using AL = AccLayout<NPOS>;
    constexpr int FW = AL::FW, G = AL::G, kProbes = AL::kProbes, kAcc = AL::kAcc;
    constexpr uint32_t MH = AL::field(0), MF = AL::field(1), MC = AL::field(2);
    constexpr uint32_t PM = (1u << NPOS) - 1u;
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
        reinterpret_cast<uint4 *>(smem + (kIdxBase - kWindowBase))[i] = reinterpret_cast<const uint4 *>(p.indexRow)[i];
    for (uint32_t i = tid; i < p.flagTabBytes / 4; i += kRowThreads) {
        const uint32_t f = reinterpret_cast<const uint32_t *>(p.flagTab)[i];   // per byte: bit0 hit, bit1 forward, bit2 canonical
        reinterpret_cast<uint32_t *>(smem + (kFlagBase - kWindowBase))[i] = ((f >> 1) & 0x03030303u) | ((f & 0x04040404u) << 5);
    }
    __syncthreads();

    const uint32_t L = p.patLen, W = p.cfg.windowSize, T = p.cfg.terminalLimit;
    const uint32_t keyMask = (1u << (2 * L)) - 1u;
    const uint32_t recHi = L | (static_cast<uint32_t>(kRegionAccounted) << 24);    // DevMatch {pos, len:16, flags:8, region:8}
    const uint32_t nUnits = p.nUnits;
    const uint32_t laneOff = lane * kRowSpan;
    const uint64_t seqBase = reinterpret_cast<uint64_t>(p.seq);
    const uint32_t slotBytes = p.slotCap * 8u;
    const bool lane0 = lane == 0;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint32_t unitSpan = static_cast<uint32_t>(kRowsPerUnit) * W;

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
    // the same start mask in accumulator layout (all three fields of a valid start)
    uint32_t fmA[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int tl = k * G + j;
            if (tl >= kProbes) continue;
#pragma unroll
            for (int o = 0; o < NPOS; ++o) {
                const int pos = tl * NPOS + o;
                if (pos < 32 && ((fm.hm >> pos) & 1u)) m |= ((1u | (1u << NPOS)) << (FW * j + o)) | (1u << (FW * j + 2 * NPOS));
            }
        }
        fmA[k] = m;
    }

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
    // One probe: the 16-bit key of the 8-mer that starts 2 * NPOS * tl bits into the lane's stream.  kRowKeyMode (build-time,
    // measured in DESIGN.md): 0 = shift + mask, both on the ALU pipe; 1 = keys inside one word are moved to the top of the
    // word by a multiply (FMA pipe) and brought down by one shift, keys that straddle two words are funnel-shifted and
    // masked; 2 = as 1, brought down by a multiply-high with 2^16 taken from a kernel parameter (so that it is not
    // strength-reduced to a shift): no ALU instruction for the in-word keys, but IMAD.HI issues at half rate.
    const uint32_t k16 = p.two16;
    auto probeKey = [&](const uint32_t (&Xv)[4], int tl) -> uint32_t {
        const int bit = 2 * NPOS * tl, wi = bit >> 5, s = bit & 31;
        if (kRowKeyMode == 0) {
            const uint32_t x = s == 0 ? Xv[wi] : (s <= 16 ? Xv[wi] >> s : __funnelshift_r(Xv[wi], Xv[wi + 1], s));
            return x & 0xFFFFu;
        }
        if (kRowKeyMode == 1) {
            if (s == 16) return Xv[wi] >> 16;
            if (s < 16) return (Xv[wi] * (1u << (16 - s))) >> 16;
            return __funnelshift_r(Xv[wi], Xv[wi + 1], s) & 0xFFFFu;
        }
        const uint32_t top = s <= 16 ? Xv[wi] * (1u << (16 - s)) : __funnelshift_r(Xv[wi], Xv[wi + 1], s) * 65536u;
        return __umulhi(top, k16);
    };
    // Records of the lane's matches `bits` (row positions, ascending) with their flags: slot entry `o` onwards, the part
    // beyond the slot in the row's overflow chunk.  Returns the sum of the flag bytes.
    auto storeRec = [&](uint8_t *slotPtr, uint32_t o, uint32_t ovBase, uint2 rec) {
        if (o < p.slotCap) reinterpret_cast<uint2 *>(slotPtr)[o] = rec;
        else if (static_cast<uint64_t>(ovBase) + (o - p.slotCap) < p.ovCap)
            reinterpret_cast<uint2 *>(p.ovArena)[ovBase + (o - p.slotCap)] = rec;
    };

    // Units are handed out by a counter (a dense unit of a telomere cap costs several sparse ones: a static split leaves
    // SMs idle at the end).  A warp always knows its next unit one unit ahead (its first row is loaded during the current
    // unit's last row) and has the index of the one after that on order.
    constexpr uint32_t kProgressBatch = 8;
    uint32_t doneRange = 0, doneCount = 0;
    uint32_t u, uNext, uPending = 0;
    if (lane0) u = atomicAdd(p.unitCursor, 2u);
    u = __shfl_sync(0xFFFFFFFFu, u, 0);
    uNext = u + 1;
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
        if (lane0) uPending = atomicAdd(p.unitCursor, 1u);
        uint64_t nextAddr = 0;                                 // the lane's first base of the first row of this warp's next unit
        uint32_t nextLim = 0;                                  // its length + 15 (lanes below it load), 0: there is none
        if (uNext < nUnits) {
            const uint4 na = __ldg(reinterpret_cast<const uint4 *>(p.units + uNext));
            nextAddr = seqBase + (static_cast<uint64_t>(na.x) | (static_cast<uint64_t>(na.y) << 32)) + laneOff;
            nextLim = min(W, na.w - na.z) + 15u;
            if (kRowPrefetch && laneOff < nextLim) {            // its first two rows: into L2 a whole unit ahead
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nextAddr));
                if (na.w - na.z > W) asm volatile("prefetch.global.L2 [%0];" ::"l"(nextAddr + W));
            }
        }
        uint8_t *slotPtr = reinterpret_cast<uint8_t *>(p.slots) + static_cast<size_t>(db.y) * slotBytes;
        uint32_t ovRows = 0;                                   // bit r: row r spilled into an overflow chunk
        // sparse unit: 16 full windows, none of them touching a terminal zone (the rule of sparseRange(), device_common.cuh)
        const bool sparse = rel > T && termEnd >= rel && termEnd - rel >= unitSpan;

        if (sparse) {
#pragma unroll 1
            for (uint32_t r = 0; r < static_cast<uint32_t>(kRowsPerUnit); ++r) {
                const uint32_t skew = static_cast<uint32_t>(laneAddr) & 15u;
                // ---- 1. pack + validate the row that has landed, then re-arm the registers with the next row -------
                uint32_t P0, P1, badBits;
                {
                    ValidAcc va;
                    P0 = pack16v(c0, va);
                    P1 = pack16v(c1, va);
                    badBits = va.bad();
                }
                const uint64_t laneAddrNext = laneAddr + W;
                {
                    uint64_t na = laneAddrNext;
                    uint32_t nl = W + 15u;
                    if (r + 1 == static_cast<uint32_t>(kRowsPerUnit)) { na = nextAddr; nl = nextLim; }
                    ldStream32If(c0, c1, na & ~static_cast<uint64_t>(15), laneOff, nl);
                    // the row after that one (inside this unit): into L2 now, so that its register load next iteration is an L2 hit
                    if (kRowPrefetch && r + 2 < static_cast<uint32_t>(kRowsPerUnit)) asm volatile("prefetch.global.L2 [%0];" ::"l"(laneAddrNext + W));
                }
                // ---- 2. remove the skew ---------------------------------------------------------------------------
                const uint32_t sh = 2u * skew;
                const uint32_t X0 = __funnelshift_r(P0, P1, sh);
                const uint32_t X1 = __funnelshift_r(P1, __shfl_down_sync(0xFFFFFFFFu, P0, 1), sh);
                const uint32_t X2 = __shfl_down_sync(0xFFFFFFFFu, X0, 1);
                // ---- 3. probes: one multiply-add each ------------------------------------------------------------
                uint32_t acc[kAcc];
#pragma unroll
                for (int k = 0; k < kAcc; ++k) acc[k] = 0;
                {
                    const uint32_t Xv[4] = {X0, X1, X2, 0u};
#pragma unroll
                    for (int tl = 0; tl < kProbes; ++tl) {
                        const uint32_t e = ldsU8Idx(probeKey(Xv, tl));
                        acc[tl / G] = e * (1u << (FW * (tl % G))) + acc[tl / G];
                    }
                }
                // ---- 4. starts whose match would cross the window end drop out (src/teloscope.cpp:439); class counts ----
                uint32_t nH = 0, nF = 0, cAny = 0;
#pragma unroll
                for (int k = 0; k < kAcc; ++k) {
                    acc[k] &= fmA[k];
                    nH += __popc(acc[k] & MH);
                    nF += __popc(acc[k] & MF);
                    cAny |= acc[k];
                }
                cAny &= MC;
                // ---- 5. A/C/G/T of the window: nE = C + G, nO = G + T, nG in 10-bit fields (a full window: <= 1002) --
                uint32_t t1 = 0;
                if (COUNT) {
                    const uint32_t Y0 = X0 & fm.m0, Y1 = X1 & fm.m1;
                    const uint32_t e = (Y0 & 0x55555555u) | ((Y1 << 1) & 0xAAAAAAAAu);    // code bit 0: C or G
                    const uint32_t o = ((Y0 >> 1) & 0x55555555u) | (Y1 & 0xAAAAAAAAu);    // code bit 1: G or T
                    t1 = __popc(e) + (__popc(o) << 10) + (__popc(e & o) << 20);
                }
                uint32_t cls = nH | (nF << 10);
                uint32_t nVrow = W;
                uint32_t info = 0;                                 // canonical matches of the row: count, or bit 31 = left to the unit epilogue
                if (__any_sync(0xFFFFFFFFu, (badBits | cAny) != 0u)) {
                const uint32_t cLanes = __ballot_sync(0xFFFFFFFFu, cAny != 0u), badLanes = __ballot_sync(0xFFFFFFFFu, badBits != 0u);
                if (badLanes == 0u && (cLanes & (cLanes - 1u)) == 0u) {
                    // ---- canonical matches in ONE lane (one window in three on random sequence holds one): the lane leaves its
                    //      registers in shared memory; the unit epilogue resolves all rows of the unit at once, one row per lane ----
                    if (cAny) {
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wb + kStashOff + r * 32), "r"(acc[0]), "r"(acc[1]), "r"(acc[2]), "r"(lane) : "memory");
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wb + kStashOff + r * 32 + 16), "r"(X0), "r"(X1), "r"(X2), "r"(0u) : "memory");
                    }
                    info = 0x80000000u;
                } else {
                    uint32_t nEmit = 0;
                    // ---- rarer: canonical matches in several lanes and / or a byte outside ACGTacgt.
                    //      Canonical starts Cpos (those on the forward strand: CF) as lane positions. ----
                    uint32_t Cpos = 0, CF = 0;
                    if (badLanes != 0u) {
                        uint32_t H = 0;                            // the lane's match starts as a position bitmap
#pragma unroll
                        for (int tl = 0; tl < kProbes; ++tl) H |= ((acc[tl / G] >> (FW * (tl % G))) & PM) << (NPOS * tl);
                        const uint4 fx = fixRow(H, X0, X1, reinterpret_cast<const uint8_t *>(laneAddr), W, L, true, COUNT);
                        H = fx.x;
                        nH = __popc(H); nF = 0;
                        uint32_t bits = H;
                        while (bits) {
                            const uint32_t t = bits & (0u - bits);
                            const uint32_t c = bfind(t);
                            bits ^= t;
                            const bool lowHalf = c < 16u;
                            const uint32_t key = __funnelshift_r(lowHalf ? X0 : X1, lowHalf ? X1 : X2, 2u * c) & keyMask;
                            const uint32_t g = ldsU8Flag(key);     // bit0 forward, bit1 canonical
                            nF += g & 1u;
                            if (g & 2u) { Cpos |= t; if (g & 1u) CF |= t; }
                        }
                        t1 = (fx.y & 0xFFFFu) + ((fx.y >> 16) << 10) + ((fx.z & 0xFFFFu) << 20);
                        nVrow = __reduce_add_sync(0xFFFFFFFFu, fx.z >> 16);
                        cls = nH | (nF << 10);
                    } else if (cAny) {                             // divergent: the (usually single) lane that holds canonical starts
#pragma unroll
                        for (int k = 0; k < kAcc; ++k) {
                            uint32_t m = acc[k] & MC;              // probes with a canonical match among their NPOS starts
                            while (m) {
                                const uint32_t b = bfind(m & (0u - m));
                                m &= m - 1u;
                                const uint32_t j = b / FW;
                                const uint32_t pos0 = (k * G + j) * NPOS;
                                uint32_t hits = (acc[k] >> (j * FW)) & PM;       // the valid starts of that probe: which one is it?
                                while (hits) {
                                    const uint32_t cpos = pos0 + bfind(hits & (0u - hits));
                                    hits &= hits - 1u;
                                    const bool lowHalf = cpos < 16u;
                                    const uint32_t key = __funnelshift_r(lowHalf ? X0 : X1, lowHalf ? X1 : X2, 2u * cpos) & keyMask;
                                    const uint32_t g = ldsU8Flag(key);           // bit0 forward, bit1 canonical
                                    if (g & 2u) { Cpos |= 1u << cpos; if (g & 1u) CF |= 1u << cpos; }
                                }
                            }
                        }
                    }
                    const uint32_t nC = __popc(Cpos);
                    uint32_t incl;
                    if (!__any_sync(0xFFFFFFFFu, nC > 1u)) {       // at most one per lane: the prefix is a ballot
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, nC != 0u);
                        incl = __popc(bal & laneLt) + nC;
                        nEmit = __popc(bal);
                    } else {
                        incl = warpInclusiveScan(nC, lane);
                        nEmit = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    }
                    uint32_t ovBase = kNoOverflow;
                    if (nEmit > p.slotCap) {
                        if (lane0) ovBase = takeOverflowChunk(p.ovCursor, p.ovChunk);
                        ovBase = __shfl_sync(0xFFFFFFFFu, ovBase, 0);
                        if (lane0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wb + kOvOff + r * 4), "r"(ovBase) : "memory");
                        ovRows |= 1u << r;
                    }
                    uint32_t o = incl - nC, bits = Cpos;
                    const uint32_t relLane = rel + laneOff;
                    while (bits) {
                        const uint32_t t = bits & (0u - bits);
                        const uint32_t c = bfind(t);
                        bits ^= t;
                        const uint32_t fl = static_cast<uint32_t>(kMatchCan) | ((CF & t) ? static_cast<uint32_t>(kMatchFwd) : 0u);
                        storeRec(slotPtr, o++, ovBase, make_uint2(relLane + c, fl * 65536u + recHi));
                    }
                    info = nEmit;
                }
                }
                const uint32_t zMask = __ballot_sync(0xFFFFFFFFu, nH == 0u);
                uint32_t sumA = 0;
                if (COUNT) sumA = __reduce_add_sync(0xFFFFFFFFu, t1);
                const uint32_t sumC = __reduce_add_sync(0xFFFFFFFFu, cls) | (nVrow << 20);
                // ---- the row's results wait in shared memory for the unit epilogue (warp-uniform values, one wavefront) ----
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wb + kRecOff + r * 16), "r"(sumA), "r"(sumC), "r"(zMask), "r"(info) : "memory");
                laneAddr = laneAddrNext; rel += W;
                slotPtr += slotBytes;
            }
        } else {
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

                // ---- 3. one table entry per NPOS start positions (its H field) ---------------------------------
                uint32_t H = 0;
                {
                    const uint32_t Xv[4] = {X0, X1, X2, 0u};
#pragma unroll
                    for (int tl = 0; tl < kProbes; ++tl) {
                        const uint32_t e = ldsU8Idx(probeKey(Xv, tl)) & PM;
                        H = e * (1u << (NPOS * tl)) + H;              // disjoint bit fields: add == or (bits >= 32 drop out)
                    }
                }
                // ---- 4. matches never cross the window end (src/teloscope.cpp:439) --------------------------------
                H &= fm.hm;

                // ---- 5. A/C/G/T of the window: packed partial sums (nE = C + G, nO = G + T, nG, bases) -------------
                uint32_t t1 = 0, t2 = fm.nv;
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
                const uint32_t sum1 = COUNT ? __reduce_add_sync(0xFFFFFFFFu, t1) : 0u;
                const uint32_t sum2 = __reduce_add_sync(0xFFFFFFFFu, t2);

                // ---- 6. emission: every lane writes its own matches, in order, behind those of the lanes before it.
                //      The match's 2L-bit key comes out of the lane's registers, the flag byte g (bit0 forward, bit1
                //      canonical, bit7 canonical) goes into the record as it is and is summed per lane: a lane has at
                //      most 32 matches, so forward + 2 canonical < 128 stays below the canonical count at bit 7. ----------
                const uint32_t rowEnd = rel + min(W, remaining);
                const bool rowHasTerm = rel <= T || rowEnd > termEnd;                               // warp-uniform
                const bool rowAllTerm = rowEnd <= T + 1u || rel >= termEnd;                         // warp-uniform
                const uint32_t nH = __popc(H);
                const uint32_t zMask = __ballot_sync(0xFFFFFFFFu, nH == 0u);
                const uint32_t incl = warpInclusiveScan(nH, lane);
                const uint32_t wMatches = __shfl_sync(0xFFFFFFFFu, incl, 31);
                const uint32_t relLane = rel + laneOff;
                uint32_t acc = 0, rowNct = 0;
                uint32_t ovBase = kNoOverflow;
                if (wMatches > p.slotCap) {
                    if (lane0) ovBase = takeOverflowChunk(p.ovCursor, p.ovChunk);
                    ovBase = __shfl_sync(0xFFFFFFFFu, ovBase, 0);
                    if (lane0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wb + kOvOff + r * 4), "r"(ovBase) : "memory");
                    ovRows |= 1u << r;
                }
                {
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
                        const bool term = rowAllTerm || (rowHasTerm && (pos <= T || pos >= termEnd));
                        storeRec(slotPtr, o++, ovBase, make_uint2(pos, g * 65536u + (term ? static_cast<uint32_t>(kMatchTerm) << 16 : 0u) + recHi));
                        acc += g;
                        if (term && !(g & 2u)) ++rowNct;
                    }
                    rowNct = __reduce_add_sync(0xFFFFFFFFu, rowNct);
                }
                // forward + 2 canonical | canonical << 16, summed over the lanes
                const uint32_t flagSum = __reduce_add_sync(0xFFFFFFFFu, (acc & 127u) | ((acc >> 7) << 16));
                const uint32_t can = flagSum >> 16, fwd = (flagSum & 0xFFFFu) - 2u * can;

                // ---- 7. the row's results wait in shared memory for the unit epilogue (the sparse rows' format) -------------
                {
                    const uint32_t sumA = (sum1 & 0xFFFFu) | ((sum1 >> 16) << 10) | ((sum2 & 0xFFFFu) << 20);
                    const uint32_t sumC = wMatches | (fwd << 10) | ((sum2 >> 16) << 20);
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(wb + kRecOff + r * 16), "r"(sumA), "r"(sumC), "r"(zMask), "r"(can) : "memory");
                    if (lane0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wb + kNctOff + r * 4), "r"(rowNct) : "memory");
                }
                laneAddr += W; rel += W;
                slotPtr += slotBytes;
            }
        }
        // ---- unit epilogue: lane r writes row r's window counters, zero-lane mask and slot bookkeeping;
        //      per-tile and per-item totals as a handful of atomics ------------------------------------------------
        __syncwarp();
        uint32_t k1 = 0, k2 = 0, k3 = 0;                       // all | canonical << 16, forward | terminal non-canonical << 16, records
        if (lane < nRows) {
            uint32_t sA, sC, zM, info;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(sA), "=r"(sC), "=r"(zM), "=r"(info) : "r"(wb + kRecOff + lane * 16));
            uint32_t ov = kNoOverflow;
            if ((ovRows >> lane) & 1u) ov = lds32(wb + kOvOff + lane * 4);
            const uint32_t nct = sparse ? 0u : lds32(wb + kNctOff + lane * 4);
            uint32_t can = info & 0x7FFFFFFFu;
            if (info >> 31) {
                // the canonical matches of this row sit in one lane of the scan: its accumulators and packed bases are in the
                // stash.  Every row of the unit is resolved here at once (this lane = this row): records in position order.
                uint32_t a[kAcc], src, Y0, Y1, Y2, pad;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(src) : "r"(wb + kStashOff + lane * 32));
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(Y0), "=r"(Y1), "=r"(Y2), "=r"(pad) : "r"(wb + kStashOff + lane * 32 + 16));
                uint2 *slot = reinterpret_cast<uint2 *>(p.slots) + static_cast<size_t>(db.y + lane) * p.slotCap;
                const uint32_t relSrc = da.z + lane * W + src * kRowSpan;
                can = 0;
#pragma unroll
                for (int k = 0; k < kAcc; ++k) {
                    uint32_t m = a[k] & MC;                        // probes with a canonical match among their NPOS starts
                    while (m) {
                        const uint32_t b = bfind(m & (0u - m));
                        m &= m - 1u;
                        const uint32_t j = b / FW;
                        const uint32_t pos0 = (k * G + j) * NPOS;
                        const uint32_t field = a[k] >> (j * FW);
                        uint32_t hits = field & PM;                // the valid starts of that probe
                        if ((hits & (hits - 1u)) == 0u) {          // one: it is the canonical one, its strand is in the F field
                            const uint32_t fl = static_cast<uint32_t>(kMatchCan) | (((field >> NPOS) & hits) ? static_cast<uint32_t>(kMatchFwd) : 0u);
                            if (hits) slot[can++] = make_uint2(relSrc + pos0 + bfind(hits), fl * 65536u + recHi);
                            continue;
                        }
                        while (hits) {
                            const uint32_t cpos = pos0 + bfind(hits & (0u - hits));
                            hits &= hits - 1u;
                            const bool lowHalf = cpos < 16u;
                            const uint32_t key = __funnelshift_r(lowHalf ? Y0 : Y1, lowHalf ? Y1 : Y2, 2u * cpos) & keyMask;
                            const uint32_t g = ldsU8Flag(key);     // bit0 forward, bit1 canonical
                            if (g & 2u) slot[can++] = make_uint2(relSrc + cpos, (g & 3u) * 65536u + recHi);
                        }
                    }
                }
            }
            const uint32_t all = sC & 1023u, fwd = (sC >> 10) & 1023u, nV = sC >> 20;
            const uint32_t emitted = sparse ? can : all;
            const uint32_t nE = sA & 1023u, nO = (sA >> 10) & 1023u, nG = sA >> 20;
            const uint32_t cA = COUNT ? nV - nE - nO + nG : 0u, cC = COUNT ? nE - nG : 0u, cG = COUNT ? nG : 0u, cT = COUNT ? nO - nG : 0u;   // zeros without TLS_OUT_ACGT
            if (p.windows8) {                                   // TLS_OUT_PACKED: counts in 8 bytes (tls_window8)
                p.windows8[db.x + lane] = packWindow8(db.x + lane, cA, cC, cG, cT, can, all, fwd, p.winExc, p.winExcCount);
            } else if (p.windows16) {                           // TLS_OUT_COMPACT: the 16-byte record at once (tls_window16)
                p.windows16[db.x + lane] = make_uint4(cA | (cC << 16), cG | (cT << 16), (can * L) | (((all - can) * L) << 16),
                                                      (fwd * L) | (((all - fwd) * L) << 16));
            } else {
                uint4 *wp = p.windows + 2 * static_cast<size_t>(db.x + lane);
                wp[0] = make_uint4(cA, cC, cG, cT);                                   // A, C, G, T
                wp[1] = make_uint4(can * L, (all - can) * L, fwd * L, (all - fwd) * L);
            }
            p.zeroLanes[db.x + lane] = zM;
            if (emitted) p.slotMeta[db.y + lane] = make_uint4(emitted, ov, can, nct);   // the table was zeroed before the launch
            k1 = all | (can << 16);
            k2 = fwd | (nct << 16);
            k3 = emitted;
        }
        __syncwarp();                                          // the rows' shared-memory results have been read: the next unit may overwrite them
        k1 = __reduce_add_sync(0xFFFFFFFFu, k1);               // <= 32 rows x 1002 per half: no carry between the halves
        k2 = __reduce_add_sync(0xFFFFFFFFu, k2);
        k3 = __reduce_add_sync(0xFFFFFFFFu, k3);
        const uint32_t uAll = k1 & 0xFFFFu, uCan = k1 >> 16, uFwd = k2 & 0xFFFFu, uNct = k2 >> 16, uRec = k3;
        if (uAll) {
            const uint32_t tile = db.y / kTileRows;
            if (lane < 3) {
                const uint32_t v = lane == 0 ? uRec : lane == 1 ? uCan : uNct;
                if (v) atomicAdd(p.tileTot + static_cast<size_t>(lane) * p.nTilesAll + tile, v);
            } else if (lane < 9) {
                uint32_t *ic = reinterpret_cast<uint32_t *>(p.itemCounts + db.w);     // all, fwd, rev, can, nonCanTerm, records
                const uint32_t v = lane == 3 ? uAll : lane == 4 ? uFwd : lane == 5 ? uAll - uFwd : lane == 6 ? uCan : lane == 7 ? uNct : uRec;
                if (v) atomicAdd(ic + (lane - 3), v);
            }
        }
        if (p.unitsPerPart) {
            // Progress for the copy stream (it waits on these words): finished units are counted per range of units.  The
            // fence that makes a unit's windows visible before its count is dear (it drains the warp's outstanding
            // memory operations), so a warp reports several units at once: when it moves on to another range, every
            // kProgressBatch units, and when it leaves.
            const uint32_t range = u / p.unitsPerPart;
            if (doneCount && (range != doneRange || doneCount >= kProgressBatch)) {
                __threadfence();
                __syncwarp();
                if (lane0) atomicAdd(p.partDone + doneRange, doneCount);
                doneCount = 0;
            }
            doneRange = range;
            ++doneCount;
            if (uNext >= nUnits) {
                __threadfence();
                __syncwarp();
                if (lane0) atomicAdd(p.partDone + doneRange, doneCount);
            }
        }
        if (uNext >= nUnits) break;
        u = uNext;
        uNext = __shfl_sync(0xFFFFFFFFu, uPending, 0);
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

// The window-row kernel and the read scorer address their shared-memory tables with absolute addresses (immediates): the
// dynamic block must start kWindowBase bytes into the CTA's shared window.  Probed once per context with the kernels' own
// launch shape; when a driver lays it out differently the context falls back to the tile kernels instead of failing later.
__global__ void __launch_bounds__(kRowThreads, 1) rowLayoutProbeKernel(unsigned int *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0) *out = smemAddr(smem);
}
cudaError_t probeRowLayout(unsigned int *dWord, bool &ok, cudaStream_t stream) {
    ok = false;
    cudaError_t e = cudaFuncSetAttribute(rowLayoutProbeKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kRowSmemTotal));
    if (e != cudaSuccess) return e;
    rowLayoutProbeKernel<<<1, kRowThreads, kRowSmemTotal, stream>>>(dWord);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    unsigned int h = 0;
    if ((e = cudaMemcpyAsync(&h, dWord, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return e;
    ok = h == kWindowBase;
    return cudaSuccess;
}

uint32_t rowNpos(uint32_t patLen) { return patLen >= 7 ? 2u : 3u; }

bool rowConfigOk(const DevConfig &c) {
    return c.mode == 0 && c.nClasses == 1 && c.cls[0].direct != 0 && c.cls[0].len <= 7 && c.cls[0].len + c.npos <= 9 &&
           c.windowSize >= kRowMinWindow && c.windowSize <= static_cast<uint32_t>(kRowMaxLen) && c.step == c.windowSize;
}

cudaError_t launchRowScan(const RowParams &p, int grid, cudaStream_t stream) {
    const size_t smem = rowSmemBytes();
    const bool count = p.cfg.countAcgt != 0;
    void (*k)(const RowParams) = nullptr;
    if (rowNpos(p.patLen) == 2) k = count ? rowScanKernel<2, true> : rowScanKernel<2, false>;
    else k = count ? rowScanKernel<3, true> : rowScanKernel<3, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k<<<grid, kRowThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tls
