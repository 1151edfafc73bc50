// Launchers of the CUDA kernels (scan_kernel.cu, block_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "device_types.hpp"

namespace tls {

size_t scanSmemBytes();
cudaError_t launchScan(const ScanParams &p, int grid, cudaStream_t stream);

// lean kernel (scan_fast.cu): genome mode, one direct pattern length, window >= kFastMinWindow
size_t fastSmemBytes();
bool fastConfigOk(const DevConfig &c);
cudaError_t launchFastScan(const FastParams &p, int grid, cudaStream_t stream);

// window-row kernel (scan_rows.cu): genome mode, one direct pattern length, kRowMinWindow <= window <= kRowMaxLen
size_t rowSmemBytes();
bool rowConfigOk(const DevConfig &c);
cudaError_t launchRowScan(const RowParams &p, int grid, cudaStream_t stream);
// fills units[0, nUnits) from the per-contig table (regions sorted by unit0)
cudaError_t launchRowPlan(const DevRowRegion *regions, uint32_t nRegions, uint32_t windowSize, DevRowUnit *units,
                          uint32_t nUnits, cudaStream_t stream);

// exclusive scans of the per-tile totals (3 arrays: tileBase[k][nTiles] = total), then ordered copy
// hostOut: 5 u64 in mapped pinned host memory: ovCursor, totals[3], redo count | status << 32
cudaError_t launchTileScan(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *tileBase, const unsigned int *ovCursor,
                           const unsigned int *redo, unsigned long long *hostOut, cudaStream_t s);
cudaError_t launchGather(const GatherParams &g, cudaStream_t s);

cudaError_t launchTrimCr(const uint8_t *seq, DevRegion *regions, uint32_t nRegions, uint32_t *itemLen, cudaStream_t s);
cudaError_t launchItemOffsets(const DevItemCounts *counts, uint32_t n, unsigned long long *itemBegin, cudaStream_t s);
// step < window: the whole window table from the sequence and the ordered match list (regions: one per non-empty
// contig, ascending winBase)
cudaError_t launchOverlapWindows(const uint8_t *seq, const DevRegion *regions, uint32_t nRegions, const DevMatch *matches,
                                 const unsigned long long *itemBegin, uint32_t nWindows, const DevConfig &cfg, uint32_t *windows,
                                 cudaStream_t s);
cudaError_t launchTerminal(const BlockParams &p, cudaStream_t s);
// runStarts: runCap entries; runCount: device counter, zeroed by the caller
cudaError_t launchIts(const BlockParams &p, const uint32_t *canIdx, unsigned long long nCan, unsigned long long *runStarts,
                      unsigned int *runCount, uint32_t runCap, cudaStream_t s);
// out receives n final 16-byte tls_match records (absPos of the item added) for the records idxList points at
cudaError_t launchSelectFinalize(const DevMatch *m, const uint32_t *idxList, unsigned long long n,
                                 const unsigned long long *itemBegin, const unsigned long long *absPos, uint32_t nItems,
                                 void *out, cudaStream_t s);

// TLS_OUT_COMPACT: 8-byte match records (item-relative position) and 16-byte window records
cudaError_t launchSelectFinalize8(const DevMatch *m, const uint32_t *idxList, unsigned long long n, void *out, cudaStream_t s);
cudaError_t launchPackWindows(const uint32_t *windows, void *windows16, unsigned long long n, cudaStream_t s);

}  // namespace tls
