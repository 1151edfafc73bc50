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
uint32_t rowNpos(uint32_t patLen);
// does the dynamic shared memory block start where the window-row kernel / read scorer expect it (absolute addressing)?
cudaError_t probeRowLayout(unsigned int *dWord, bool &ok, cudaStream_t stream);       // start offsets one table byte answers: 3 (pattern length <= 6), 2 (length 7)
cudaError_t launchRowScan(const RowParams &p, int grid, cudaStream_t stream);
// fills units[0, nUnits) from the per-contig table (regions sorted by unit0)
cudaError_t launchRowPlan(const DevRowRegion *regions, uint32_t nRegions, uint32_t windowSize, DevRowUnit *units,
                          uint32_t nUnits, cudaStream_t stream);

// read scorer (scan_reads.cu): ReadTelomereFilter::matches fused, one warp per read; one direct pattern length, maxMatchDist >= 31
size_t readSmemBytes();
bool readConfigOk(const DevConfig &c);
cudaError_t launchReadScan(const ReadParams &p, int grid, cudaStream_t stream);

// exclusive scans of the per-tile totals (3 arrays: tileBase[k][nTiles] = total), then ordered copy
// hostOut: 5 u64 in mapped pinned host memory: ovCursor, totals[3], redo count | status << 32
// blockSums: scratch of tileScanScratchBytes(nTiles) bytes
size_t tileScanScratchBytes(uint32_t nTiles);
cudaError_t launchTileScan(const uint32_t *tileTot, uint32_t nTiles, unsigned long long *tileBase, unsigned long long *blockSums,
                           const unsigned int *ovCursor, const unsigned int *redo, unsigned long long *hostOut, uint32_t *denseTiles,
                           unsigned int *denseCount, cudaStream_t s);
cudaError_t launchGather(const GatherParams &g, cudaStream_t s, cudaStream_t denseStream);   // sparse tiles on s, dense tiles (listed by the tile scan) on denseStream

cudaError_t launchTrimCr(const uint8_t *seq, DevRegion *regions, uint32_t nRegions, uint32_t *itemLen, cudaStream_t s);
cudaError_t launchTrimCrItems(const uint8_t *seq, const DevItemInfo *items, uint32_t n, uint32_t *itemLen, cudaStream_t s);
cudaError_t launchItemOffsets(const DevItemCounts *counts, const DevItemInfo *items, uint32_t n, unsigned long long *itemBegin, cudaStream_t s);
// step < window: the whole window table from the sequence and the ordered match list (regions: one per non-empty
// contig, ascending winBase)
cudaError_t launchOverlapWindows(const uint8_t *seq, const DevRegion *regions, uint32_t nRegions, const DevMatch *matches,
                                 const unsigned long long *itemBegin, uint32_t nWindows, const DevConfig &cfg, uint32_t *windows,
                                 cudaStream_t s);
cudaError_t launchTerminal(const BlockParams &p, cudaStream_t s);
// runStarts: runCap entries; runCount: device counter, zeroed by the caller
cudaError_t launchIts(const BlockParams &p, const uint32_t *canIdx, unsigned long long nCan, unsigned long long *runStarts,
                      unsigned int *runCount, uint32_t runCap, cudaStream_t s);
// tiles[index[k]] = compact[k]: the tiles of the lean / general kernels go to their places in the batch's tile table
cudaError_t launchScatterTiles(const DevTile *compact, const uint32_t *index, uint32_t n, DevTile *tiles, cudaStream_t s);
// interstitial blocks of window-row items (conn: nCan + 32 bytes; chainStarts: chainCap entries; chainCount: zeroed by the caller)
cudaError_t launchItsRows(const BlockParams &p, const uint32_t *canIdx, unsigned long long nCan, uint8_t *conn,
                          unsigned long long *chainStarts, unsigned int *chainCount, uint32_t chainCap, cudaStream_t s);
// windows [begin, end) of the ranges (rangePrefix: nRanges + 1 exclusive prefix of their lengths): zero the 32-byte
// counters / pack them into tls_window16 records
cudaError_t launchZeroWindowRanges(uint32_t *windows, const uint2 *ranges, const uint32_t *rangePrefix, uint32_t nRanges,
                                   uint32_t total, cudaStream_t s);
cudaError_t launchPackWindowRanges(const uint32_t *windows, void *windows16, const uint2 *ranges, const uint32_t *rangePrefix,
                                   uint32_t nRanges, uint32_t total, cudaStream_t s);
// out receives n final 16-byte tls_match records (absPos of the item added) for the records idxList points at
cudaError_t launchSelectFinalize(const DevMatch *m, const uint32_t *idxList, unsigned long long n,
                                 const unsigned long long *itemBegin, const unsigned long long *absPos, uint32_t nItems,
                                 void *out, cudaStream_t s);

// TLS_OUT_COMPACT: 8-byte match records (item-relative position) and 16-byte window records
cudaError_t launchSelectFinalize8(const DevMatch *m, const uint32_t *idxList, unsigned long long n, void *out, cudaStream_t s);
cudaError_t launchPackWindows(const uint32_t *windows, void *windows16, unsigned long long n, cudaStream_t s);
// TLS_OUT_PACKED: 8-byte window records (counts, pattern length L; exceptions for 255 matches or more) and 4-byte match records
cudaError_t launchPackWindows8(const uint32_t *windows, void *windows8, unsigned long long n, uint32_t L, void *winExc, unsigned int *winExcCount,
                               cudaStream_t s);
cudaError_t launchPackWindow8Ranges(const uint32_t *windows, void *windows8, const uint2 *ranges, const uint32_t *rangePrefix, uint32_t nRanges,
                                    uint32_t total, uint32_t L, void *winExc, unsigned int *winExcCount, cudaStream_t s);

}  // namespace tls
