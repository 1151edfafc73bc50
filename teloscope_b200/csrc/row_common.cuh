// Device helpers shared by the window-row scan kernel (scan_rows.cu) and the read scorer (scan_reads.cu): streaming
// loads of a lane's 32 bases, SWAR pack to 2-bit codes with the exact "every byte is one of ACGTacgt" test, and the
// exact out-of-line path of a row that is short or holds other bytes.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "device_common.cuh"

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

// Slow exact path of a row that holds a byte outside ACGTacgt (IUPAC letters, '\r', ...): per-lane
// bitmap of such bases in row coordinates and the counts of the lane's valid bases, in the packed form
// (x: invalid bitmap, y: nE | nO << 16, z: nG | nValid << 16 with nE = C + G, nO = G + T).  The reference's Trie
// stops at such a byte (include/teloscope.h:27-35,63-66) and analyzeWindow does not count it (src/teloscope.cpp:418-435).
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
    uint32_t t1 = 0, t2 = nv << 16;
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
}  // namespace tls
