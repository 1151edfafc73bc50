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

// exclusive scan of per-tile match totals (tileBase[nTiles] = total), then ordered copy
cudaError_t launchTileScan(const GatherParams &g, cudaStream_t s);
cudaError_t launchGather(const GatherParams &g, cudaStream_t s);

cudaError_t launchTrimCr(const uint8_t *seq, DevRegion *regions, uint32_t nRegions, uint32_t *itemLen, cudaStream_t s);
cudaError_t launchItemOffsets(const DevItemCounts *counts, uint32_t n, unsigned long long *itemBegin, cudaStream_t s);
cudaError_t launchTerminal(const BlockParams &p, cudaStream_t s);
cudaError_t launchIts(const BlockParams &p, unsigned long long nMatches, cudaStream_t s);

uint32_t selectChunks(unsigned long long nMatches);
// chunkScratch: 2 * selectChunks(n) u32; totals: 2 u64 (canonical, terminal non-canonical);
// outCan / outNon receive final 16-byte tls_match records (absPos of the item added)
cudaError_t launchSelect(const DevMatch *m, unsigned long long n, uint32_t *chunkScratch, unsigned long long *totals,
                         const unsigned long long *itemBegin, const unsigned long long *absPos, uint32_t nItems,
                         void *outCan, uint32_t capCan, void *outNon, uint32_t capNon, cudaStream_t s);

}  // namespace tls
