// Read scorer for sm_100a: ReadTelomereFilter::matches (src/read-filter.cpp:10-45) fused into one kernel, ONE WARP PER READ.
//
// The reference scans a read like a contig end with terminalLimit = UINT32_MAX / 2 (the whole read is "in the terminal
// zone", so every run may start anywhere), builds the terminal blocks of both strands (getTerminalBlocks,
// src/teloscope.cpp:29-176) and keeps the read when either strand yields a block.  Nothing but that one bit leaves the
// device here: no match records, no arena, no gather, no block records.
//
//   * a warp takes the next read from a work counter and streams it in rows of kReadStride start positions (a lane owns
//     32 bases, the same register pipeline of 16-byte streaming loads, SWAR pack and exact validity test as the window-row
//     kernel, row_common.cuh); matches may straddle rows, so a row loads patLen - 1 bases more than it owns starts;
//   * 8-mer byte table -> bitmap of match starts per lane; the flag table splits it by strand and marks the canonical ones;
//   * per strand the block state machine runs on LANES, not on matches: with maxMatchDist >= 31 all matches inside one
//     lane (32 consecutive start positions) belong to one run, so a lane is one step (first / last position, count,
//     canonical count).  Rows without a canonical match on that strand cannot complete a kept run (Phase 1 needs
//     canonicalCount > 0, src/teloscope.cpp:76-77): they only update the open run (start, last, count) with a few ballots;
//   * Phase 1 filter with the reference's float32 compare, Phase 2 merge on the fly; a read is decided (and left) as soon as
//     a merged block reaches minBlockLen.
// Configurations outside this kernel's scope (several pattern lengths, patterns longer than 7, maxMatchDist < 31) keep using
// the general scan + block kernels.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_common.cuh"
#include "kernels.hpp"
#include "row_common.cuh"

namespace tls {

namespace {

constexpr uint32_t kWindowBase = 0x400;
constexpr uint32_t kFlagBase = kWindowBase;                   // 16 KiB flag table
constexpr uint32_t kIdxBase = 0x10000u;                       // 64 KiB 8-mer table (one byte per key: NPOS hit bits)
static_assert(kFlagBase + kFastFlagBytes <= kIdxBase, "flag table must end before the 8-mer table");
constexpr uint32_t kReadSmemTotal = kIdxBase + 65536u - kWindowBase;

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

// One strand of one read: the open run (Phase 1) and the block being merged (Phase 2).  Warp-uniform.
struct Strand {
    bool active = false;        // a run is open
    uint32_t start = 0, last = 0, count = 0, can = 0;
    bool curActive = false;     // a kept run has been seen: the block under construction
    uint32_t curStart = 0, curEnd = 0;
    uint32_t total = 0;         // matches of the strand (only tracked when minBlockCounts < 2, see below)
    bool pend = false;          // some block has reached minBlockLen
};

}  // namespace

template <int NPOS>
__global__ void __launch_bounds__(kReadThreads, 1) readScanKernel(const ReadParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    if (smemAddr(smem) != kWindowBase) {
        if (tid == 0 && blockIdx.x == 0) atomicExch(p.status, 1u);
        return;
    }
    for (int i = tid; i < 65536 / 16; i += kReadThreads)
        reinterpret_cast<uint4 *>(smem + (kIdxBase - kWindowBase))[i] = reinterpret_cast<const uint4 *>(p.index8)[i];
    for (uint32_t i = tid; i < p.flagTabBytes / 4; i += kReadThreads)      // per byte: bit0 hit, bit1 forward, bit2 canonical
        reinterpret_cast<uint32_t *>(smem + (kFlagBase - kWindowBase))[i] = reinterpret_cast<const uint32_t *>(p.flagTab)[i];
    __syncthreads();

    const DevConfig &c = p.cfg;
    const uint32_t L = p.patLen, K = c.maxMatchDist;
    const uint32_t keyMask = (1u << (2 * L)) - 1u;
    const uint32_t laneOff = lane * kRowSpan;
    const uint32_t laneLt = (1u << lane) - 1u;
    const uint64_t seqBase = reinterpret_cast<uint64_t>(p.seq);
    const bool trackTotal = c.minBlockCounts < 2u;          // otherwise a kept run already has two matches (src/teloscope.cpp:646,649)
    constexpr int kProbes = (kRowSpan + NPOS - 1) / NPOS;

    // Phase 1 filter (src/teloscope.cpp:75-92) and Phase 2 merge (:130-174) for a run that has ended
    auto closeRun = [&](Strand &s) {
        const uint32_t end = s.last + L;
        const float rhs = __fmul_rn(c.minBlockDensity, __ull2float_rn(static_cast<unsigned long long>(end - s.start)));
        if (s.count >= c.minBlockCounts && s.can > 0 && __uint2float_rn(s.can * L) >= rhs) {
            if (s.curActive && s.start - s.curEnd <= c.maxBlockDist) s.curEnd = end;
            else { s.curStart = s.start; s.curEnd = end; s.curActive = true; }
            if (s.curEnd - s.curStart >= c.minBlockLen) s.pend = true;
        }
        s.active = false;
    };
    // The matches M of one strand in this row (Cs: the canonical ones), lane by lane in position order
    auto feed = [&](Strand &s, uint32_t M, uint32_t Cs, uint32_t rowPos) {
        const uint32_t Bm = __ballot_sync(0xFFFFFFFFu, M != 0u);
        if (Bm == 0u) return;
        const uint32_t nM = __popc(M);
        if (trackTotal) s.total += __reduce_add_sync(0xFFFFFFFFu, nM);
        const uint32_t firstPos = rowPos + laneOff + (__ffs(M) - 1), lastPos = rowPos + laneOff + (31 - __clz(M));   // garbage when M == 0
        const bool anyCan = __any_sync(0xFFFFFFFFu, Cs != 0u);
        if (!anyCan && !(s.active && s.can > 0u)) {
            // No canonical match here and none in the open run: whatever run ends inside this row is dropped by Phase 1.
            // Only the run that is open at the end of the row matters: its start, count and last match.
            const uint32_t lower = Bm & laneLt;
            const uint32_t prevLast = __shfl_sync(0xFFFFFFFFu, lastPos, lower ? 31 - __clz(lower) : 0);
            const bool hasPrev = lower != 0u || s.active;
            const uint32_t prev = lower ? prevLast : s.last;
            const bool isStart = M != 0u && (!hasPrev || firstPos - prev > K);     // the lane's first match opens a new run
            const uint32_t Bs = __ballot_sync(0xFFFFFFFFu, isStart);
            const uint32_t newLast = __shfl_sync(0xFFFFFFFFu, lastPos, 31 - __clz(Bm));
            if (Bs == 0u) {
                s.count += __reduce_add_sync(0xFFFFFFFFu, nM);
            } else {
                const uint32_t hl = 31 - __clz(Bs);
                s.start = __shfl_sync(0xFFFFFFFFu, firstPos, hl);
                s.count = __reduce_add_sync(0xFFFFFFFFu, lane >= hl ? nM : 0u);
                s.can = 0; s.active = true;
            }
            s.last = newLast;
            return;
        }
        // Canonical content: the lanes with matches fall into SEGMENTS at the run starts (a lane's first match out of reach of
        // the match before it).  The lanes before the first start continue the open run; a start closes the run before it; a
        // closed run without a canonical match is dropped by Phase 1, so only the segments with canonical matches and the last
        // one (it stays open) are summed up, with two reductions and two shuffles each.  A telomeric row is one segment.
        const uint32_t lower = Bm & laneLt;
        const uint32_t prevLast = __shfl_sync(0xFFFFFFFFu, lastPos, lower ? 31 - __clz(lower) : 0);
        const bool isStart = M != 0u && (lower ? firstPos - prevLast > K : (!s.active || firstPos - s.last > K));
        const uint32_t Bs = __ballot_sync(0xFFFFFFFFu, isStart);
        const uint32_t canCnt = __popc(Cs);
        auto below32 = [](int k) { return k >= 32 ? 0xFFFFFFFFu : (1u << k) - 1u; };      // lanes before lane k
        auto segment = [&](uint32_t seg, bool fresh) {
            const bool inSeg = (seg >> lane) & 1u;
            const uint32_t cnt = __reduce_add_sync(0xFFFFFFFFu, inSeg ? nM : 0u), can = __reduce_add_sync(0xFFFFFFFFu, inSeg ? canCnt : 0u);
            const uint32_t la = __shfl_sync(0xFFFFFFFFu, lastPos, 31 - __clz(seg));
            if (fresh) {
                s.active = true; s.start = __shfl_sync(0xFFFFFFFFu, firstPos, __ffs(seg) - 1); s.count = cnt; s.can = can;
            } else {
                s.count += cnt; s.can += can;
            }
            s.last = la;
        };
        const int firstStart = Bs ? __ffs(Bs) - 1 : 32;
        const uint32_t seg0 = Bm & below32(firstStart);
        if (seg0) segment(seg0, false);                               // the open run goes on (the first lane is no start)
        if (Bs) {
            if (s.active) closeRun(s);                                // the run before the first start is over
            const int lastStart = 31 - __clz(Bs);
            uint32_t cm = __ballot_sync(0xFFFFFFFFu, Cs != 0u) & ~below32(firstStart) & below32(lastStart);   // canonical lanes of closed middle segments
            while (cm) {
                const int b = __ffs(cm) - 1;
                const int h = 31 - __clz(Bs & below32(b + 1));                            // the start of b's segment
                const int nx = __ffs(Bs & ~below32(b + 1)) - 1;                           // the next start: b < lastStart
                segment(Bm & ~below32(h) & below32(nx), true);
                closeRun(s);
                cm &= ~below32(nx);
            }
            segment(Bm & ~below32(lastStart), true);                  // the last segment stays open
        }
        // a run that no start of a later row can reach any more is closed now: the next rows then take the cheap path above
        if (s.active && rowPos + kReadStride - s.last > K) closeRun(s);
    };

    uint4 c0 = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u), c1 = c0;
    // row j of a read of n bases owns the starts [j * kReadStride, ...) and loads L - 1 bases more
    auto rowBases = [&](uint32_t n, uint32_t rs) { return min(n - rs, static_cast<uint32_t>(kReadStride) + L - 1u); };

    unsigned long long matchTotal = 0;
    // Reads are handed out by a counter.  A warp knows its next read one read ahead (the first row of that read is loaded
    // during the last row of the current one) and has the index of the one after that on order.
    uint32_t r = 0, rNext, rPending = 0;
    if (lane == 0) r = atomicAdd(p.cursor, 2u);
    r = __shfl_sync(0xFFFFFFFFu, r, 0);
    rNext = r + 1;
    bool primed = false;                                      // c0 / c1 hold (or are receiving) row 0 of read r
    while (r < p.nItems) {
        const uint32_t n = p.itemLen[r];
        const uint64_t readAddr = seqBase + p.items[r].bufStart;
        if (lane == 0) rPending = atomicAdd(p.cursor, 1u);
        uint32_t nNext = 0;
        uint64_t nextAddr = 0;
        if (rNext < p.nItems) { nNext = p.itemLen[rNext]; nextAddr = seqBase + p.items[rNext].bufStart + laneOff; }
        if (!primed && n >= L) ldStream32If(c0, c1, (readAddr + laneOff) & ~static_cast<uint64_t>(15), laneOff, rowBases(n, 0) + 15u);
        primed = false;
        Strand fw, rv;
        uint32_t nHits = 0;
        bool decided = false;
        if (n >= L) {
            for (uint32_t rs = 0; rs + L <= n; rs += kReadStride) {
                const uint64_t laneAddr = readAddr + rs + laneOff;
                const uint32_t skew = static_cast<uint32_t>(laneAddr) & 15u;
                const uint32_t nb = rowBases(n, rs);
                uint32_t P0, P1, badBits;
                {
                    ValidAcc va;
                    P0 = pack16v(c0, va);
                    P1 = pack16v(c1, va);
                    badBits = va.bad();
                }
                {   // re-arm: the next row of this read, or the first row of the warp's next read
                    const uint32_t rsN = rs + kReadStride;
                    const bool more = rsN + L <= n;
                    const uint64_t na = more ? laneAddr + kReadStride : nextAddr;
                    const uint32_t nl = more ? rowBases(n, rsN) + 15u : (nNext >= L ? rowBases(nNext, 0) + 15u : 0u);
                    ldStream32If(c0, c1, na & ~static_cast<uint64_t>(15), laneOff, nl);
                    primed = !more && nNext >= L;
                    // the row after that one: into L2 now
                    if (kRowPrefetch && rsN + kReadStride + laneOff < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(laneAddr + 2 * kReadStride));
                }
                const uint32_t sh = 2u * skew;
                const uint32_t X0 = __funnelshift_r(P0, P1, sh);
                const uint32_t X1 = __funnelshift_r(P1, __shfl_down_sync(0xFFFFFFFFu, P0, 1), sh);
                const uint32_t X2 = __shfl_down_sync(0xFFFFFFFFu, X0, 1);
                const bool anyBad = __any_sync(0xFFFFFFFFu, badBits != 0);
                uint32_t H = 0;
                {
                    const uint32_t Xv[4] = {X0, X1, X2, 0u};
#pragma unroll
                    for (int tl = 0; tl < kProbes; ++tl) {
                        const int bit = 2 * NPOS * tl, wi = bit >> 5, s = bit & 31;
                        const uint32_t x = s == 0 ? Xv[wi] : (s <= 16 ? Xv[wi] >> s : __funnelshift_r(Xv[wi], Xv[wi + 1], s));
                        const uint32_t e = ldsU8Idx(x & 0xFFFFu);
                        H = e * (1u << (NPOS * tl)) + H;
                    }
                }
                // starts of this row whose match ends inside the read
                H &= bmskClamp(max(static_cast<int>(nb) - static_cast<int>(laneOff) - static_cast<int>(L) + 1, 0));
                if (anyBad) H = fixRow(H, X0, X1, reinterpret_cast<const uint8_t *>(laneAddr), nb, L, true, false).x;
                // strand and canonical flag of every match
                uint32_t MF = 0, CB = 0, bits = H;
                while (bits) {
                    const uint32_t t = bits & (0u - bits);
                    const uint32_t cpos = bfind(t);
                    bits ^= t;
                    const bool lowHalf = cpos < 16u;
                    const uint32_t key = __funnelshift_r(lowHalf ? X0 : X1, lowHalf ? X1 : X2, 2u * cpos) & keyMask;
                    const uint32_t g = ldsU8Flag(key);
                    if (g & 2u) MF |= t;
                    if (g & 4u) CB |= t;
                }
                nHits += __popc(H);
                feed(fw, H & MF, CB & MF, rs);
                feed(rv, H & ~MF, CB & ~MF, rs);
                if (fw.pend | rv.pend) {
                    if ((fw.pend && (!trackTotal || fw.total >= 2u)) || (rv.pend && (!trackTotal || rv.total >= 2u))) { decided = true; break; }
                }
            }
        }
        bool pass = decided;
        if (!decided) {
            if (fw.active) closeRun(fw);
            if (rv.active) closeRun(rv);
            pass = (fw.pend && (!trackTotal || fw.total >= 2u)) || (rv.pend && (!trackTotal || rv.total >= 2u));
        }
        if (lane == 0) p.pass[r] = pass ? 1 : 0;
        matchTotal += __reduce_add_sync(0xFFFFFFFFu, nHits);       // warp-uniform
        if (decided) primed = false;                               // left early: the registers hold a row of this read
        r = rNext;
        rNext = __shfl_sync(0xFFFFFFFFu, rPending, 0);
    }
    if (lane == 0 && matchTotal) atomicAdd(p.matchTotal, matchTotal);
}

size_t readSmemBytes() { return kReadSmemTotal; }

bool readConfigOk(const DevConfig &c) {
    return c.nClasses == 1 && c.cls[0].direct != 0 && c.cls[0].len <= 7 && c.cls[0].len + c.npos <= 9 && c.maxMatchDist >= 31u;
}

cudaError_t launchReadScan(const ReadParams &p, int grid, cudaStream_t stream) {
    const size_t smem = readSmemBytes();
    void (*k)(const ReadParams) = nullptr;
    switch (p.cfg.npos) {
        case 2: k = readScanKernel<2>; break;
        case 3: k = readScanKernel<3>; break;
        default: k = readScanKernel<4>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    k<<<grid, kReadThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace tls
