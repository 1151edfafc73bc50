// Kernels that turn the per-row match slots into what SegmentData holds besides the window counters:
//   tileBlockSumsKernel / tileScanKernel   exclusive prefixes over tiles of (all, canonical, terminal non-canonical) records,
//                       the list of dense tiles, the batch's totals published to mapped host memory
//   gatherKernel / gatherDenseKernel        slot arena -> one ordered match list + the -m selections (index lists; with
//                       TLS_OUT_PACKED the final 4-byte records): a warp per tile of sparse units, a block per dense tile
//   itemOffsetsKernel   exclusive scan of per-item record counts
//   terminalKernel      Teloscope::getTerminalBlocks (src/teloscope.cpp:29-176), one warp per (item, arm), record groups
//                       cut into segments at the run starts, Phase 2 merged on the fly
//   itsSeedKernel / itsRunKernel            Teloscope::getInterstitialBlocks (src/teloscope.cpp:179-256) for items with
//                       complete lists: a block needs >= 4 canonical matches, so only runs that hold them are walked
//   itsConnKernel / itsChainKernel / itsWalkKernel   the same for window-row items (sparse lists): chains of canonical matches
//                       the zero-lane masks cannot prove apart are settled on the sequence
//   selectFinalizeKernel / selectFinalize8Kernel     final 16- / 8-byte records of the -m lists
//   overlapWindowsKernel, pack / zero / scatter helpers
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <type_traits>

#include "device_common.cuh"
#include "row_common.cuh"
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
__global__ void __launch_bounds__(1024) itemOffsetsKernel(const DevItemCounts *counts, const DevItemInfo *items, uint32_t n,
                                                          unsigned long long *itemBegin) {
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        // records of the item in the ordered list: every match, or what the window-row kernel emitted (sparse units: canonical only)
        const unsigned long long v = i < n ? ((items && items[i].rowItem) ? counts[i].records : counts[i].all) : 0;
        const unsigned long long excl = blockExclusive1024(v, warpSums, &carry);
        if (i < n) itemBegin[i] = excl;
    }
    if (threadIdx.x == 0) itemBegin[n] = carry;
}

// ------------------------------------------------------------------------------------------------
// Per-tile totals (written by the scan kernels with atomics) -> exclusive prefixes.
// tileTot / tileBase hold three arrays back to back: match records, canonical, terminal non-canonical.
// Two launches over blocks of kScanChunk tiles: block sums, then every block adds up the sums of the blocks
// before it and scans its own tiles.
// The totals, the overflow cursor and the redo counters are also published to `hostOut` (mapped
// pinned host memory: [0] ovCursor, [1..3] totals, [4] redo count | status << 32), so that the host
// needs no copy-engine transfer (which would queue behind the window counters going home).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kDenseTile = 128;                                 // a tile with more records than this goes to the block-per-tile gather kernel
constexpr uint32_t kScanChunk = 4096;                               // tiles per block: 1024 threads x 4

__global__ void __launch_bounds__(1024) tileBlockSumsKernel(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *blockSums,
                                                            uint32_t *denseTiles, unsigned int *denseCount) {
    __shared__ unsigned long long warpSums[32];
    const uint32_t k = blockIdx.y, b = blockIdx.x, nBlk = gridDim.x;
    const uint32_t *in = tileTot + static_cast<size_t>(k) * nTiles;
    const uint32_t i0 = b * kScanChunk + threadIdx.x * 4;
    unsigned long long v = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (i0 + j < nTiles) {
            const uint32_t t = in[i0 + j];
            v += t;
            // the tiles the block-per-tile gather kernel takes (in no particular order: every record's place is computed)
            if (k == 0 && denseTiles && t > kDenseTile) denseTiles[atomicAdd(denseCount, 1u)] = i0 + j;
        }
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = warpSums[threadIdx.x];
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        if (threadIdx.x == 0) blockSums[static_cast<size_t>(k) * nBlk + b] = v;
    }
}

__global__ void __launch_bounds__(1024) tileScanKernel(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *tileBase,
                                                       const unsigned long long *blockSums, const unsigned int *ovCursor,
                                                       const unsigned int *redo, unsigned long long *hostOut) {
    __shared__ unsigned long long warpSums[32];
    __shared__ unsigned long long carry;
    const uint32_t k = blockIdx.y, b = blockIdx.x, nBlk = gridDim.x;
    // offset of this block: the sums of the blocks before it (and, in the last block, the grand total)
    unsigned long long before = 0, all = 0;
    for (uint32_t i = threadIdx.x; i < nBlk; i += 1024) {
        const unsigned long long v = blockSums[static_cast<size_t>(k) * nBlk + i];
        all += v;
        if (i < b) before += v;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { before += __shfl_xor_sync(0xFFFFFFFFu, before, d); all += __shfl_xor_sync(0xFFFFFFFFu, all, d); }
    __shared__ unsigned long long redBefore[32], redAll[32];
    if ((threadIdx.x & 31) == 0) { redBefore[threadIdx.x >> 5] = before; redAll[threadIdx.x >> 5] = all; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long sb = 0, sa = 0;
        for (int i = 0; i < 32; ++i) { sb += redBefore[i]; sa += redAll[i]; }
        carry = sb;
        if (b == nBlk - 1) {
            tileBase[static_cast<size_t>(k) * (nTiles + 1) + nTiles] = sa;
            hostOut[1 + k] = sa;
            if (k == 0) {
                hostOut[0] = *reinterpret_cast<const unsigned long long *>(ovCursor);
                hostOut[4] = redo ? (static_cast<unsigned long long>(redo[0]) | (static_cast<unsigned long long>(redo[1]) << 32)) : 0ull;
            }
            __threadfence_system();
        }
    }
    __syncthreads();
    const uint32_t *in = tileTot + static_cast<size_t>(k) * nTiles;
    unsigned long long *out = tileBase + static_cast<size_t>(k) * (nTiles + 1);
    const uint32_t i0 = b * kScanChunk + threadIdx.x * 4;
    uint32_t v[4];
    unsigned long long sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = i0 + j < nTiles ? in[i0 + j] : 0u; sum += v[j]; }
    unsigned long long excl = blockExclusive1024(sum, warpSums, &carry);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < nTiles) out[i0 + j] = excl;
        excl += v[j];
    }
}

// ------------------------------------------------------------------------------------------------
// Slot arena -> ordered match list.  gatherKernel: one WARP per tile (32 row slots) of sparse units, which holds a handful
// of canonical records: lane r copies row r's records itself; persistent warps request the next tile's header ahead.
// gatherDenseKernel: a tile with more than kDenseTile records (dense units, other scan kernels; listed by the tile scan)
// is copied by one BLOCK, warp w taking four slots with coalesced 8-byte accesses.  (All window / item accounting was done
// by the scan kernels.)
// ------------------------------------------------------------------------------------------------
constexpr int kGatherThreads = 256;
#ifndef TLS_GATHER_BLOCKS_PER_SM
#define TLS_GATHER_BLOCKS_PER_SM 16
#endif
constexpr uint32_t kGatherBlocksPerSm = TLS_GATHER_BLOCKS_PER_SM;   // 4 x 256 threads are resident; several short waves balance better than one long one (measured: 4: 135 us, 16: 108 us)

struct TileHeader {
    unsigned long long base, baseEnd, baseCan, baseNct;
    uint4 meta;                                                     // records, overflow base, canonical, terminal non-canonical
};
// every load of a tile's header is issued together (one round trip to memory, not three)
__device__ __forceinline__ TileHeader loadTileHeader(const GatherParams &g, uint32_t tile, uint32_t lane) {
    TileHeader h;
    h.base = g.tileBase[tile];
    h.baseEnd = g.tileBase[tile + 1];
    h.baseCan = g.tileBase[static_cast<size_t>(g.nTiles) + 1 + tile];
    h.baseNct = g.tileBase[2 * (static_cast<size_t>(g.nTiles) + 1) + tile];
    h.meta = g.slotMeta[static_cast<size_t>(tile) * kTileRows + lane];
    return h;
}

struct TilePrefix { uint32_t exAll, exCan, exNct; };
// exclusive prefixes over the 32 row slots (a tile holds at most 32 x ovChunk records: 16 bits per field would not do)
__device__ __forceinline__ TilePrefix tilePrefix(const uint4 &meta, uint32_t lane) {
    uint32_t inclAll = meta.x, inclCan = meta.z, inclNct = meta.w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, inclAll, d), c = __shfl_up_sync(0xFFFFFFFFu, inclCan, d),
                       n = __shfl_up_sync(0xFFFFFFFFu, inclNct, d);
        if (lane >= static_cast<uint32_t>(d)) { inclAll += a; inclCan += c; inclNct += n; }
    }
    return TilePrefix{inclAll - meta.x, inclCan - meta.z, inclNct - meta.w};
}


// A tile of sparse units: lane r copies row r's records.  Tiles with many records are only listed here.
__device__ __forceinline__ void gatherTile(const GatherParams &g, uint32_t tile, uint32_t lane, const TileHeader &hd) {
    const size_t slot0 = static_cast<size_t>(tile) * kTileRows;
    const unsigned long long base = hd.base, baseCan = hd.baseCan, baseNct = hd.baseNct;
    const uint4 meta = hd.meta;
    const uint32_t total = static_cast<uint32_t>(hd.baseEnd - base);
    if (total == 0) return;
    if (total > kDenseTile) return;                                 // listed by tileBlockSumsKernel for gatherDenseKernel
    const TilePrefix px = tilePrefix(meta, lane);
    const uint32_t exAll = px.exAll, exCan = px.exCan, exNct = px.exNct;
    const uint2 *slots = reinterpret_cast<const uint2 *>(g.slots);
    const uint2 *ovArena = reinterpret_cast<const uint2 *>(g.ovArena);
    uint2 *out = reinterpret_cast<uint2 *>(g.matches);
    // lane r: row r's records, one after the other
    const uint2 *src = slots + (slot0 + lane) * g.slotCap;
    unsigned long long o = base + exAll, oc = baseCan + exCan, on = baseNct + exNct;
    auto put = [&](const uint2 rec) {
        if (o < g.matchCap) out[o] = make_uint2(rec.x, rec.y & 0x00FFFFFFu);
        const uint32_t flags = (rec.y >> 16) & 0xFFu;
        const uint32_t m4 = rec.x | ((flags & kMatchFwd) ? 0x80000000u : 0u);             // tls_match4
        if (flags & kMatchCan) {
            if (oc < g.selCap) { g.canIdx[oc] = static_cast<uint32_t>(o); if (g.sel4Can) g.sel4Can[oc] = m4; }
            ++oc;
        } else if (flags & kMatchTerm) {
            if (on < g.selCapNct) { if (g.sel4Nct) g.sel4Nct[on] = m4; else g.nctIdx[on] = static_cast<uint32_t>(o); }
            ++on;
        }
        ++o;
    };
    // the first records of the row (a sparse row holds one or two) are all requested before any is stored
    constexpr uint32_t kAhead = 4;
    const uint32_t direct = min(min(meta.x, g.slotCap), kAhead);
    uint2 first[kAhead];
#pragma unroll
    for (uint32_t k = 0; k < kAhead; ++k) first[k] = k < direct ? __ldg(src + k) : make_uint2(0u, 0u);
#pragma unroll
    for (uint32_t k = 0; k < kAhead; ++k) if (k < direct) put(first[k]);
    for (uint32_t k = direct; k < meta.x; ++k) {
        uint2 rec;
        if (k < g.slotCap) rec = __ldg(src + k);
        else if (static_cast<unsigned long long>(meta.y) + (k - g.slotCap) < g.ovCap) rec = ovArena[meta.y + (k - g.slotCap)];
        else { ++o; continue; }                                 // overflow arena exhausted: the host reruns the scan
        put(rec);
    }
}

// A tile with many records: one BLOCK per tile, warp w copies the slots [4 w, 4 w + 4) with coalesced 8-byte accesses, the first
// 32 records of its four slots requested together.
__global__ void __launch_bounds__(kGatherThreads) gatherDenseKernel(const GatherParams g) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t nDense = *g.denseCount;
    const uint2 *slots = reinterpret_cast<const uint2 *>(g.slots);
    const uint2 *ovArena = reinterpret_cast<const uint2 *>(g.ovArena);
    uint2 *out = reinterpret_cast<uint2 *>(g.matches);
    const uint32_t laneLt = (1u << lane) - 1u;
    constexpr int kSlotBatch = kTileRows / (kGatherThreads / 32);   // 4
    for (uint32_t i = blockIdx.x; i < nDense; i += gridDim.x) {
        const uint32_t tile = g.denseTiles[i];
        const TileHeader hd = loadTileHeader(g, tile, lane);
        const size_t slot0 = static_cast<size_t>(tile) * kTileRows;
        const unsigned long long base = hd.base, baseCan = hd.baseCan, baseNct = hd.baseNct;
        const uint4 meta = hd.meta;
        const TilePrefix px = tilePrefix(meta, lane);
        const uint32_t exAll = px.exAll, exCan = px.exCan, exNct = px.exNct;
        const int s0 = static_cast<int>(warp) * kSlotBatch;
        uint2 head[kSlotBatch];
#pragma unroll
        for (int u = 0; u < kSlotBatch; ++u) {
            const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, meta.x, s0 + u);
            head[u] = lane < min(cnt, g.slotCap) ? __ldg(slots + (slot0 + s0 + u) * g.slotCap + lane) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < kSlotBatch; ++u) {
            const int s = s0 + u;
            const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, meta.x, s);
            if (cnt == 0) continue;
            const unsigned long long ov = __shfl_sync(0xFFFFFFFFu, meta.y, s);
            const bool anySel = __shfl_sync(0xFFFFFFFFu, meta.z | meta.w, s) != 0;
            const unsigned long long o = base + __shfl_sync(0xFFFFFFFFu, exAll, s);
            unsigned long long oc = baseCan + __shfl_sync(0xFFFFFFFFu, exCan, s), on = baseNct + __shfl_sync(0xFFFFFFFFu, exNct, s);
            const uint2 *src = slots + (slot0 + s) * g.slotCap;
            for (uint32_t k0 = 0; k0 < cnt; k0 += 32) {
                const uint32_t k = k0 + lane;
                uint2 rec = head[u];
                bool have = k < cnt;
                if (have && k0 != 0) {
                    if (k < g.slotCap) rec = __ldg(src + k);
                    else if (ov + (k - g.slotCap) < g.ovCap) rec = ovArena[ov + (k - g.slotCap)];
                    else have = false;                          // overflow arena exhausted: the host reruns the scan
                } else if (have && k >= g.slotCap) {            // first chunk of a slot with fewer than 32 places
                    if (ov + (k - g.slotCap) < g.ovCap) rec = ovArena[ov + (k - g.slotCap)];
                    else have = false;
                }
                if (have && o + k < g.matchCap) out[o + k] = make_uint2(rec.x, rec.y & 0x00FFFFFFu);
                if (!anySel) continue;
                // index lists of the -m selections, in genome order
                const uint32_t flags = (rec.y >> 16) & 0xFFu;
                const bool isC = have && (flags & kMatchCan);
                const bool isN = have && (flags & (kMatchCan | kMatchTerm)) == kMatchTerm;
                const uint32_t bc = __ballot_sync(0xFFFFFFFFu, isC), bn = __ballot_sync(0xFFFFFFFFu, isN);
                const uint32_t m4 = rec.x | ((flags & kMatchFwd) ? 0x80000000u : 0u);         // tls_match4
                if (isC && oc + __popc(bc & laneLt) < g.selCap) {
                    g.canIdx[oc + __popc(bc & laneLt)] = static_cast<uint32_t>(o + k);
                    if (g.sel4Can) g.sel4Can[oc + __popc(bc & laneLt)] = m4;
                }
                if (isN && on + __popc(bn & laneLt) < g.selCapNct) {
                    if (g.sel4Nct) g.sel4Nct[on + __popc(bn & laneLt)] = m4;
                    else g.nctIdx[on + __popc(bn & laneLt)] = static_cast<uint32_t>(o + k);
                }
                oc += __popc(bc);
                on += __popc(bn);
            }
        }
    }
}

// Persistent warps: a warp walks tiles with a stride and requests the next tile's header before it copies the current one.
__global__ void __launch_bounds__(kGatherThreads) gatherKernel(const GatherParams g) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * (kGatherThreads / 32);
    uint32_t tile = blockIdx.x * (kGatherThreads / 32) + (threadIdx.x >> 5);
    if (tile >= g.nTiles) return;
    TileHeader hd = loadTileHeader(g, tile, lane);
    for (;;) {
        const uint32_t next = tile + stride;
        TileHeader hn = hd;
        if (next < g.nTiles) hn = loadTileHeader(g, next, lane);
        gatherTile(g, tile, lane, hd);
        if (next >= g.nTiles) break;
        tile = next;
        hd = hn;
    }
}

__global__ void __launch_bounds__(256) scatterTilesKernel(const DevTile *compact, const uint32_t *index, uint32_t n, DevTile *tiles) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) tiles[index[k]] = compact[k];
}
cudaError_t launchScatterTiles(const DevTile *compact, const uint32_t *index, uint32_t n, DevTile *tiles, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    scatterTilesKernel<<<(n + 255) / 256, 256, 0, s>>>(compact, index, n, tiles);
    return cudaGetLastError();
}

cudaError_t launchTileScan(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *tileBase, unsigned long long *blockSums,
                           const unsigned int *ovCursor, const unsigned int *redo, unsigned long long *hostOut, uint32_t *denseTiles,
                           unsigned int *denseCount, cudaStream_t s) {
    const uint32_t nBlk = std::max<uint32_t>(1u, (nTiles + kScanChunk - 1) / kScanChunk);
    tileBlockSumsKernel<<<dim3(nBlk, 3), 1024, 0, s>>>(tileTot, nTiles, blockSums, denseTiles, denseCount);
    tileScanKernel<<<dim3(nBlk, 3), 1024, 0, s>>>(tileTot, nTiles, tileBase, blockSums, ovCursor, redo, hostOut);
    return cudaGetLastError();
}
size_t tileScanScratchBytes(uint32_t nTiles) { return static_cast<size_t>(std::max<uint32_t>(1u, (nTiles + kScanChunk - 1) / kScanChunk)) * 3 * 8; }
cudaError_t launchGather(const GatherParams &g, cudaStream_t s, cudaStream_t denseStream) {
    if (g.nTiles == 0) return cudaSuccess;
    const uint32_t blocks = std::min<uint32_t>((g.nTiles + kGatherThreads / 32 - 1) / (kGatherThreads / 32), 148u * kGatherBlocksPerSm);
    gatherKernel<<<blocks, kGatherThreads, 0, s>>>(g);
    gatherDenseKernel<<<std::min<uint32_t>(g.nTiles, 148u * 4u), kGatherThreads, 0, denseStream>>>(g);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Terminal blocks
// ------------------------------------------------------------------------------------------------
namespace {

// The match that starts at one position of a window-row item, straight from the sequence: what the scan kernel found there
// (same table, same rules: a match never crosses its window's end, src/teloscope.cpp:439; a byte outside ACGTacgt stops
// the Trie walk, include/teloscope.h:27-35).  The walks below use it where the record list is sparse.
struct Rescan {
    const uint8_t *seq;        // the item's first base
    const uint8_t *flagTab;    // 4^L bytes: bit0 hit, bit1 forward, bit2 canonical
    uint32_t n, W, L;
    const DevConfig *cfg;
    // The bases of the pattern that would start at pos: unconditional loads (the batch is readable past its end), so a
    // caller can have several positions in flight before it looks at any of them.
    __device__ void fetch(uint32_t pos, uint32_t (&ch)[7]) const {
        const uint8_t *s = seq + (pos < n ? pos : 0u);
#pragma unroll
        for (uint32_t j = 0; j < 7; ++j) ch[j] = j < L ? __ldg(s + j) : static_cast<uint32_t>('A');   // L <= 7
    }
    __device__ uint32_t resolve(uint32_t pos, const uint32_t (&ch)[7]) const {
        if (pos >= n) return 0u;
        const uint32_t wStart = divWindow(*cfg, pos) * W;
        const uint32_t wLen = min(W, n - wStart);
        if (pos - wStart + L > wLen) return 0u;
        // four bases per word: the scan kernels' exact ACGTacgt test and 2-bit packing (row_common.cuh); bases past L are 'A' = code 0
        const uint32_t w0 = ch[0] | (ch[1] << 8) | (ch[2] << 16) | (ch[3] << 24), w1 = ch[4] | (ch[5] << 8) | (ch[6] << 16) | (0x41u << 24);
        ValidAcc v;
        const uint32_t key = (packWord(w0, v) >> 24) | ((packWord(w1, v) >> 24) << 8);
        return v.bad() == 0u ? __ldg(flagTab + key) : 0u;
    }
    __device__ uint32_t at(uint32_t pos) const {
        uint32_t ch[7];
        fetch(pos, ch);
        return resolve(pos, ch);
    }
};

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
        // A window-row item lists every match only outside [spLo, spHi) (sparseRange): the walk reads the list up to that
        // edge and, if a run is still open there, goes on over the sequence itself (below).
        uint32_t spLo = 0xFFFFFFFFu, spHi = 0xFFFFFFFFu;
        if (p.items && p.items[item].rowItem) sparseRange(n, c.terminalLimit, c.windowSize, spLo, spHi);
        const bool hasSparse = spLo != 0xFFFFFFFFu;
        bool atEdge = false;
        // lane l holds the l-th next record in walk order; the groups of 32 behind the current one are requested kAhead groups
        // ahead and only looked at (the sparse edge) when their turn comes
        auto load = [&](long long off, uint2 &rec, bool &valid) {
            const long long i = fromStart ? begin + off + lane : end - 1 - off - lane;
            valid = off < end - begin && (fromStart ? i < end : i >= begin);
            rec = valid ? __ldg(recs + i) : make_uint2(0u, 0u);
        };
        constexpr int kAhead = 3;
        uint2 qRec[kAhead];
        bool qValid[kAhead];
#pragma unroll
        for (int k = 0; k < kAhead; ++k) load(32ll * k, qRec[k], qValid[k]);
        for (long long off = 0; off < end - begin && !done; off += 32) {
            const uint2 rec = qRec[0];
            bool valid = qValid[0];
#pragma unroll
            for (int k = 0; k + 1 < kAhead; ++k) { qRec[k] = qRec[k + 1]; qValid[k] = qValid[k + 1]; }
            load(off + 32ll * kAhead, qRec[kAhead - 1], qValid[kAhead - 1]);
            if (valid && hasSparse && (fromStart ? rec.x >= spLo : rec.x < spHi)) { valid = false; atEdge = true; }
            if (__any_sync(0xFFFFFFFFu, atEdge) && !__any_sync(0xFFFFFFFFu, valid)) break;     // only records beyond the edge are left
            const uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid && ((rec.y >> 16) & kMatchFwd) == want);
            if (!mask) continue;
            // The group's records of the wanted strand fall into SEGMENTS: a segment starts where the gap to the previous
            // record exceeds maxMatchDist (a run start, src/teloscope.cpp:98-121).  The records before the first start
            // continue the open run; every start closes the run before it.  A closed run without a canonical match fails the
            // Phase 1 filter (:76-77), so only segments with canonical content and the group's last segment (it stays open)
            // need their sums, which come from ballots and warp reductions; the first run start outside the terminal zone
            // ends the walk.  Telomere arrays are one segment per group, random sequence a handful of cheap ones.
            const bool mine = (mask >> lane) & 1u;
            const uint32_t len = rec.y & 0xFFFFu;
            const bool can = mine && ((rec.y >> 16) & kMatchCan);
            const uint32_t canM = __ballot_sync(0xFFFFFFFFu, can);
            const uint32_t below = mask & ((1u << lane) - 1u);
            uint32_t prevPos = __shfl_sync(0xFFFFFFFFu, rec.x, below ? 31 - __clz(below) : 0);
            if (!below) prevPos = run.prev;
            const uint32_t gapL = fromStart ? rec.x - prevPos : prevPos - rec.x;
            const bool isStart = mine && ((!below && !inRun) || gapL > c.maxMatchDist);
            uint32_t st = __ballot_sync(0xFFFFFFFFu, isStart);
            const uint32_t outZ = __ballot_sync(0xFFFFFFFFu, isStart && !inZone(rec.x));
            uint32_t live = mask;
            if (outZ) { const uint32_t lim = (1u << (__ffs(outZ) - 1)) - 1u; live &= lim; st &= lim; }
            auto takeSegment = [&](uint32_t seg, bool fresh) {
                const bool inSeg = (seg >> lane) & 1u;
                const uint32_t cntM = __popc(seg), canCnt = __popc(seg & canM);
                const uint32_t lenSum = __reduce_add_sync(0xFFFFFFFFu, inSeg ? len : 0u);
                const uint32_t canLen = canCnt ? __reduce_add_sync(0xFFFFFFFFu, (inSeg && can) ? len : 0u) : 0u;
                const int last = 31 - __clz(seg);
                const uint32_t lastPos = __shfl_sync(0xFFFFFFFFu, rec.x, last), lastLen = __shfl_sync(0xFFFFFFFFu, len, last);
                if (fresh) {
                    const int first = __ffs(seg) - 1;
                    const uint32_t firstPos = __shfl_sync(0xFFFFFFFFu, rec.x, first), firstLen = __shfl_sync(0xFFFFFFFFu, len, first);
                    run.n = cntM; run.nFwd = want ? cntM : 0u; run.nCan = canCnt;
                    run.cov = lenSum; run.fwdCov = want ? lenSum : 0u; run.canCov = canLen;
                    if (fromStart) { run.start = firstPos; run.end = lastPos + lastLen; }
                    else { run.end = firstPos + firstLen; run.start = lastPos; }
                } else {
                    run.n += cntM; run.nFwd += want ? cntM : 0u; run.nCan += canCnt;
                    run.cov += lenSum; run.fwdCov += want ? lenSum : 0u; run.canCov += canLen;
                    if (fromStart) run.end = lastPos + lastLen; else run.start = lastPos;
                }
                run.prev = lastPos;
            };
            auto below32 = [](int k) { return k >= 32 ? 0xFFFFFFFFu : (1u << k) - 1u; };      // lanes before lane k
            if (live) {
                const int firstStart = st ? __ffs(st) - 1 : 32;
                const uint32_t seg0 = live & below32(firstStart);
                if (seg0) takeSegment(seg0, false);                   // the open run goes on (the first record is no start)
                if (st) {
                    if (inRun && run.nCan > 0u) feedSub(run);         // the run before the first start is over
                    inRun = true;
                    const int lastStart = 31 - __clz(st);
                    // segments that begin and end inside the group: only those with a canonical match can pass Phase 1
                    uint32_t cm = canM & live & ~below32(firstStart) & below32(lastStart);
                    while (cm) {
                        const int b = __ffs(cm) - 1;
                        const int h = 31 - __clz(st & below32(b + 1));                    // the start of b's segment
                        const int nx = __ffs(st & ~below32(b + 1)) - 1;                   // the next start: b < lastStart
                        takeSegment(live & ~below32(h) & below32(nx), true);
                        feedSub(run);
                        cm &= ~below32(nx);
                    }
                    takeSegment(live & ~below32(lastStart), true);    // the last segment stays open
                }
            }
            if (outZ) {                                               // a run start outside the zone: the walk ends (:98-102,117-121)
                if (inRun && run.nCan > 0u) feedSub(run);
                inRun = false;
                done = true;
            }
        }
        if (hasSparse && inRun && !done) {
            // The open run may go on inside the sparse part: same-strand matches from the sequence, 32 positions per step,
            // until the run breaks.  Every position there is outside the terminal zone, so the first match that does not
            // continue the run would start a run out of zone, which ends the walk (src/teloscope.cpp:98-102,117-121).
            Rescan rs{p.seq + p.items[item].bufStart, p.flagTab, n, c.windowSize, p.patLen, &c};
            for (long long base = fromStart ? static_cast<long long>(spLo) : static_cast<long long>(spHi) - 1;
                 !done && (fromStart ? base < static_cast<long long>(n) : base >= 0); base += fromStart ? 32 : -32) {
                const uint32_t reach = fromStart ? static_cast<uint32_t>(base) - run.prev : run.prev - static_cast<uint32_t>(base);
                if (reach > c.maxMatchDist) break;                      // nothing in reach any more: the run is over
                const long long q = fromStart ? base + lane : base - lane;
                const uint32_t f = q >= 0 ? rs.at(static_cast<uint32_t>(q)) : 0u;
                uint32_t mask = __ballot_sync(0xFFFFFFFFu, (f & 1u) && ((f >> 1) & 1u) == (want ? 1u : 0u));
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t pos = static_cast<uint32_t>(fromStart ? base + b : base - b);
                    const uint32_t fl = want | ((__shfl_sync(0xFFFFFFFFu, f, b) & 4u) ? static_cast<uint32_t>(kMatchCan) : 0u);
                    const uint32_t gap = fromStart ? pos - run.prev : run.prev - pos;
                    if (gap > c.maxMatchDist) { done = true; break; }
                    if (fromStart) run.end = pos + p.patLen; else run.start = pos;
                    run.absorb(pos, p.patLen, fl);
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
    if (p.items && p.items[item].rowItem) return;                 // window-row items: itsConnKernel / itsWalkKernel below
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
    const int lane = threadIdx.x & 31;
    const uint32_t nRuns = min(*runCount, runCap);
    const DevConfig &c = p.cfg;
    for (uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < nRuns; gw += (gridDim.x * blockDim.x) >> 5) {
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
    if (lane != 0) continue;
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
}

// ------------------------------------------------------------------------------------------------
// Interstitial blocks of window-row items.  Their record list holds every match only outside the sparse
// part, so runs are not read from it.  A block needs 4 canonical matches in one run (src/teloscope.cpp:186),
// i.e. 4 CONSECUTIVE canonical matches that are pairwise connected by matches at most maxMatchDist apart:
//   itsConnKernel   per canonical match j: can j and j + 1 belong to one run?  "No" is proven from the scan
//                   kernel's per-window masks of 32-position lanes without a match start: ceil(maxMatchDist / 32)
//                   such lanes in a row between the two break every chain.  On random sequence that settles
//                   all but a handful of pairs without touching the sequence.
//   itsChainKernel  starts of maximal chains of >= 4 "maybe connected" canonical matches
//   itsWalkKernel   one warp per chain: the runs around its canonical matches, rebuilt exactly from the sequence
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) itsConnKernel(const BlockParams p, const uint32_t *canIdx, unsigned long long nCan, uint8_t *conn) {
    const unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= nCan) return;
    const DevConfig &c = p.cfg;
    uint8_t out = 0;
    do {
        if (j + 1 >= nCan) break;
        const unsigned long long idx = canIdx[j], idx2 = canIdx[j + 1];
        const uint2 *recs = reinterpret_cast<const uint2 *>(p.matches);
        const uint32_t a = __ldg(&recs[idx].x), b = __ldg(&recs[idx2].x);   // requested before the item search, not after it
        uint32_t lo = 0, hi = p.nItems;                               // item: last i with itemBegin[i] <= idx
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (p.itemBegin[mid] <= idx) lo = mid + 1; else hi = mid;
        }
        const uint32_t item = lo - 1;
        if (!p.items[item].rowItem || idx2 >= p.itemBegin[item + 1]) break;
        const uint32_t fwdB = p.fwdBoundary[item], revB = p.revBoundary[item];
        if (!(fwdB < revB) || p.itemCounts[item].all < 2) break;      // src/teloscope.cpp:653
        if (a < fwdB || b >= revB) break;                             // runs live in [fwdBoundary, revBoundary), :188-191,235
        out = 1;
        const uint32_t m = c.maxMatchDist == 0 ? 1u : (c.maxMatchDist + 31u) / 32u;      // lanes without a start that break a chain
        if (m > 31u) break;
        const uint32_t W = c.windowSize, n = p.itemLen[item];
        const uint32_t rowA = divWindow(c, a), rowB = divWindow(c, b);
        for (uint32_t row = rowA; row <= rowB && row < rowA + 3u; ++row) {
            const uint32_t rowStart = row * W, nFull = min(W, n - rowStart) / 32u;   // lanes that lie entirely inside the window
            if (nFull < m) continue;
            const uint32_t Z = p.zeroLanes[p.items[item].winBase + row];
            uint32_t R = Z;
            for (uint32_t i = 1; i < m; ++i) R &= Z >> i;             // bit l: lanes l .. l + m - 1 hold no start
            uint32_t lMin = 0, lEnd = nFull - m + 1u;                 // admissible first lanes [lMin, lEnd)
            if (row == rowA) lMin = (a - rowStart) / 32u + 1u;        // the stretch starts after a
            if (row == rowB) lEnd = min(lEnd, (b - rowStart) / 32u >= m ? (b - rowStart) / 32u - m + 1u : 0u);   // and ends at or before b
            if (lMin >= lEnd) continue;
            const uint32_t range = (lEnd >= 32u ? 0xFFFFFFFFu : (1u << lEnd) - 1u) & ~((1u << lMin) - 1u);
            if (R & range) { out = 0; break; }
        }
    } while (false);
    conn[j] = out;
}

__global__ void __launch_bounds__(256) itsChainKernel(const uint8_t *conn, unsigned long long nCan, unsigned long long *chainStarts,
                                                      unsigned int *chainCount, uint32_t chainCap) {
    const unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j + 3 >= nCan) return;
    if (!(conn[j] && conn[j + 1] && conn[j + 2]) || (j > 0 && conn[j - 1])) return;
    const unsigned int slot = atomicAdd(chainCount, 1u);
    if (slot < chainCap) chainStarts[slot] = j;
}

constexpr int kWalkBlocks = 4;          // positions rescanned per step of a walk: 32 x kWalkBlocks

__global__ void __launch_bounds__(128) itsWalkKernel(const BlockParams p, const uint32_t *canIdx, const uint8_t *conn,
                                                     const unsigned long long *chainStarts, const unsigned int *chainCount, uint32_t chainCap) {
    const int lane = threadIdx.x & 31;
    const uint32_t nChains = min(*chainCount, chainCap);
    const DevConfig &c = p.cfg;
    for (uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; gw < nChains; gw += (gridDim.x * blockDim.x) >> 5) {
    unsigned long long j = chainStarts[gw];
    unsigned long long jEnd = j;                                      // last canonical match of the chain (conn of the batch's last is 0,
    for (;;) {                                                        // the array is padded by 32 bytes): 32 links per step
        const uint32_t open32 = __ballot_sync(0xFFFFFFFFu, conn[jEnd + lane] != 0);
        if (open32 != 0xFFFFFFFFu) { jEnd += __ffs(~open32) - 1; break; }
        jEnd += 32;
    }
    const uint2 *recs = reinterpret_cast<const uint2 *>(p.matches);
    const unsigned long long idx0 = canIdx[j];
    uint32_t lo = 0, hi = p.nItems;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (p.itemBegin[mid] <= idx0) lo = mid + 1; else hi = mid;
    }
    const uint32_t item = lo - 1;
    const uint32_t n = p.itemLen[item], L = p.patLen, K = c.maxMatchDist;
    const uint32_t fwdB = p.fwdBoundary[item], revB = p.revBoundary[item];
    const Rescan rs{p.seq + p.items[item].bufStart, p.flagTab, n, c.windowSize, L, &c};
    // the run rightwards from a match at `from`: its matches, forward and canonical ones, and the last match's position
    auto walkRight = [&](uint32_t from, auto blocksTag, uint32_t &cnt, uint32_t &nFwd, uint32_t &nCan, uint32_t &prev) {
        constexpr int kBlocks = decltype(blocksTag)::value;           // 32 x kBlocks positions per step, all loads in flight together
        const uint32_t f0 = rs.at(from);
        cnt = 1; nFwd = (f0 >> 1) & 1u; nCan = (f0 >> 2) & 1u; prev = from;
        bool open = true;
        for (unsigned long long base = static_cast<unsigned long long>(from) + 1; open && base < revB; base += 32 * kBlocks) {
            if (base - prev > K) break;
            uint32_t f[kBlocks], ch[kBlocks][7];
#pragma unroll
            for (int blk = 0; blk < kBlocks; ++blk) rs.fetch(static_cast<uint32_t>(min(base + 32u * blk + lane, static_cast<unsigned long long>(n))), ch[blk]);
#pragma unroll
            for (int blk = 0; blk < kBlocks; ++blk) {
                const unsigned long long q = base + 32u * blk + lane;
                f[blk] = q < revB ? rs.resolve(static_cast<uint32_t>(q), ch[blk]) : 0u;
            }
#pragma unroll
            for (int blk = 0; blk < kBlocks; ++blk) {
                uint32_t mask = __ballot_sync(0xFFFFFFFFu, (f[blk] & 1u) != 0u);
                if (!open) mask = 0;
                const uint32_t fwdM = __ballot_sync(0xFFFFFFFFu, (f[blk] & 3u) == 3u), canM = __ballot_sync(0xFFFFFFFFu, (f[blk] & 5u) == 5u);
                uint32_t taken = 0;
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t pos = static_cast<uint32_t>(base + 32u * blk + b);
                    if (pos - prev > K) { open = false; break; }
                    taken |= 1u << b; prev = pos;
                }
                cnt += __popc(taken); nFwd += __popc(taken & fwdM); nCan += __popc(taken & canM);
                if (open && static_cast<unsigned long long>(base) + 32u * (blk + 1) - prev > K) open = false;   // the next block starts out of reach
            }
        }
    };
    while (j <= jEnd) {
        const uint32_t seed = recs[canIdx[j]].x;
        // The seed is the first canonical match of its run: the canonical matches before it in the chain are either in a run
        // walked already or proven out of reach.  So the run holds 4 canonical matches only if the part from the seed
        // rightwards does (src/teloscope.cpp:186): a short walk settles the chains that the masks could not break.
        uint32_t start = seed, cnt, nFwd, nCan, prev;
        walkRight(seed, std::integral_constant<int, 2>{}, cnt, nFwd, nCan, prev);
        if (nCan >= 4) {
            // ---- the run's first match: leftwards from the seed while the next match is within reach ----
            bool open = true;
            for (long long top = static_cast<long long>(seed) - 1; open && top >= static_cast<long long>(fwdB); top -= 32) {
                if (start - static_cast<uint32_t>(top) > K) break;
                const long long q = top - lane;
                const uint32_t f = q >= static_cast<long long>(fwdB) ? rs.at(static_cast<uint32_t>(q)) : 0u;
                uint32_t mask = __ballot_sync(0xFFFFFFFFu, (f & 1u) != 0u);
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t pos = static_cast<uint32_t>(top - b);
                    if (start - pos > K) { open = false; break; }
                    start = pos;
                }
            }
            // ---- the whole run, rightwards from its first match ----
            if (start != seed) walkRight(start, std::integral_constant<int, kWalkBlocks>{}, cnt, nFwd, nCan, prev);
        }
        if (lane == 0 && nCan >= 4) {
            const uint32_t len = prev + L - start;
            // computeBlockLabel, include/teloscope.h:217-222 (float32)
            const float ratio = __fdiv_rn(__fmul_rn(__uint2float_rn(nFwd), 100.0f), __uint2float_rn(cnt));
            const char label = ratio > 66.6f ? 'p' : (ratio < 33.3f ? 'q' : 'b');
            if (len >= c.itsMinLen && !(label == 'b' && nFwd < 2 && (cnt - nFwd) < 2)) {
                DevBlock b{};
                b.item = item; b.order = 0; b.start = start; b.len = len; b.counts = cnt; b.fwdCount = nFwd;
                b.canCount = nCan; b.totalCov = cnt * L; b.fwdCov = nFwd * L; b.canCov = nCan * L;
                b.validOr = 1; b.label = label; b.kind = 1;
                emitBlock(p, b);
            }
        }
        // ---- the chain's next canonical match behind this run ----
        ++j;
        for (;;) {                                                    // 32 candidates per step
            const unsigned long long jj = j + lane;
            const bool behind = jj <= jEnd && recs[canIdx[jj]].x <= prev;
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, behind);
            if (m != 0xFFFFFFFFu) { j += __ffs(~m) - 1; break; }
            j += 32;
        }
    }
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

// The windows of the ranges[k] = [x, y) (rangePrefix: exclusive prefix of their lengths, nRanges + 1 entries), one thread per
// window: the contigs the window-row kernel does not take (their counters are accumulated with atomics by the other scan
// kernels, so they start from zero and are packed afterwards; the window-row kernel writes final records itself).
__device__ __forceinline__ uint32_t windowOfRanges(const uint2 *ranges, const uint32_t *rangePrefix, uint32_t nRanges, uint32_t i) {
    uint32_t lo = 0, hi = nRanges;                                 // last range with prefix <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (rangePrefix[mid] <= i) lo = mid; else hi = mid;
    }
    return ranges[lo].x + (i - rangePrefix[lo]);
}
__global__ void __launch_bounds__(256) zeroWindowRangesKernel(uint4 *windows, const uint2 *ranges, const uint32_t *rangePrefix,
                                                              uint32_t nRanges, uint32_t total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t w = windowOfRanges(ranges, rangePrefix, nRanges, i);
    windows[2 * w] = make_uint4(0u, 0u, 0u, 0u);
    windows[2 * w + 1] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void __launch_bounds__(256) packWindowRangesKernel(const uint4 *in, uint4 *out, const uint2 *ranges, const uint32_t *rangePrefix,
                                                              uint32_t nRanges, uint32_t total) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t w = windowOfRanges(ranges, rangePrefix, nRanges, i);
    const uint4 a = in[2 * w], b = in[2 * w + 1];
    out[w] = make_uint4(a.x | (a.y << 16), a.z | (a.w << 16), b.x | (b.y << 16), b.z | (b.w << 16));
}

// tls_window (8 x u32) -> tls_window8 (counts of matches): one pattern length L, window <= 1023
__device__ __forceinline__ void storeWindow8(uint2 *out, size_t w, const uint4 a, const uint4 b, uint32_t L, uint4 *winExc, unsigned int *winExcCount) {
    out[w] = packWindow8(static_cast<uint32_t>(w), a.x, a.y, a.z, a.w, b.x / L, (b.x + b.y) / L, b.z / L, winExc, winExcCount);
}
__global__ void __launch_bounds__(256) packWindows8Kernel(const uint4 *in, uint2 *out, unsigned long long n, uint32_t L, uint4 *winExc,
                                                          unsigned int *winExcCount) {
    const unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    storeWindow8(out, i, in[2 * i], in[2 * i + 1], L, winExc, winExcCount);
}
__global__ void __launch_bounds__(256) packWindow8RangesKernel(const uint4 *in, uint2 *out, const uint2 *ranges, const uint32_t *rangePrefix,
                                                               uint32_t nRanges, uint32_t total, uint32_t L, uint4 *winExc, unsigned int *winExcCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t w = windowOfRanges(ranges, rangePrefix, nRanges, i);
    storeWindow8(out, w, in[2 * w], in[2 * w + 1], L, winExc, winExcCount);
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
cudaError_t launchZeroWindowRanges(uint32_t *windows, const uint2 *ranges, const uint32_t *rangePrefix, uint32_t nRanges,
                                   uint32_t total, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    zeroWindowRangesKernel<<<(total + 255) / 256, 256, 0, s>>>(reinterpret_cast<uint4 *>(windows), ranges, rangePrefix, nRanges, total);
    return cudaGetLastError();
}
cudaError_t launchPackWindowRanges(const uint32_t *windows, void *windows16, const uint2 *ranges, const uint32_t *rangePrefix,
                                   uint32_t nRanges, uint32_t total, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    packWindowRangesKernel<<<(total + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4 *>(windows), static_cast<uint4 *>(windows16),
                                                             ranges, rangePrefix, nRanges, total);
    return cudaGetLastError();
}
cudaError_t launchPackWindows8(const uint32_t *windows, void *windows8, unsigned long long n, uint32_t L, void *winExc, unsigned int *winExcCount,
                               cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    packWindows8Kernel<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint4 *>(windows), static_cast<uint2 *>(windows8), n, L,
                                                                                static_cast<uint4 *>(winExc), winExcCount);
    return cudaGetLastError();
}
cudaError_t launchPackWindow8Ranges(const uint32_t *windows, void *windows8, const uint2 *ranges, const uint32_t *rangePrefix, uint32_t nRanges,
                                    uint32_t total, uint32_t L, void *winExc, unsigned int *winExcCount, cudaStream_t s) {
    if (total == 0) return cudaSuccess;
    packWindow8RangesKernel<<<(total + 255) / 256, 256, 0, s>>>(reinterpret_cast<const uint4 *>(windows), static_cast<uint2 *>(windows8), ranges, rangePrefix,
                                                              nRanges, total, L, static_cast<uint4 *>(winExc), winExcCount);
    return cudaGetLastError();
}
cudaError_t launchItemOffsets(const DevItemCounts *counts, const DevItemInfo *items, uint32_t n, unsigned long long *itemBegin, cudaStream_t s) {
    itemOffsetsKernel<<<1, 1024, 0, s>>>(counts, items, n, itemBegin);
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
    const unsigned long long warps = std::min<unsigned long long>(std::min<unsigned long long>(runCap, nCan / 4 + 1), 148 * 32);
    itsRunKernel<<<static_cast<unsigned int>((warps * 32 + 127) / 128), 128, 0, s>>>(p, runStarts, runCount, runCap);
    return cudaGetLastError();
}
cudaError_t launchItsRows(const BlockParams &p, const uint32_t *canIdx, unsigned long long nCan, uint8_t *conn,
                          unsigned long long *chainStarts, unsigned int *chainCount, uint32_t chainCap, cudaStream_t s) {
    if (nCan < 4) return cudaSuccess;
    const unsigned int blocks = static_cast<unsigned int>((nCan + 255) / 256);
    itsConnKernel<<<blocks, 256, 0, s>>>(p, canIdx, nCan, conn);
    itsChainKernel<<<blocks, 256, 0, s>>>(conn, nCan, chainStarts, chainCount, chainCap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // the number of chains is only known on the device: enough warps for the capacity, surplus warps exit
    const unsigned long long warps = std::min<unsigned long long>(std::min<unsigned long long>(chainCap, nCan / 4 + 1), 148 * 32);
    itsWalkKernel<<<static_cast<unsigned int>((warps * 32 + 127) / 128), 128, 0, s>>>(p, canIdx, conn, chainStarts, chainCount, chainCap);
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

// the same for the read scorer, which works from the per-item table
__global__ void trimCrItemsKernel(const uint8_t *seq, const DevItemInfo *items, uint32_t n, uint32_t *itemLen) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t len = itemLen[i];
    if (len && seq[items[i].bufStart + len - 1] == '\r') itemLen[i] = len - 1;
}
cudaError_t launchTrimCrItems(const uint8_t *seq, const DevItemInfo *items, uint32_t n, uint32_t *itemLen, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    trimCrItemsKernel<<<(n + 255) / 256, 256, 0, s>>>(seq, items, n, itemLen);
    return cudaGetLastError();
}

cudaError_t launchTrimCr(const uint8_t *seq, DevRegion *regions, uint32_t nRegions, uint32_t *itemLen, cudaStream_t s) {
    if (nRegions == 0) return cudaSuccess;
    trimCrKernel<<<(nRegions + 255) / 256, 256, 0, s>>>(seq, regions, nRegions, itemLen);
    return cudaGetLastError();
}

}  // namespace tls
