// Plain structs shared between the host side of the library (api.cu) and the kernels.
#pragma once

#include <cstdint>

#include <vector_types.h>

namespace tls {

// ---- geometry of the fused scan kernel ------------------------------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kSpanBytes = 48;                    // bases per lane per row (3 x 16-byte chunks)
constexpr int kRowBytes = 32 * kSpanBytes;        // 1536
constexpr int kRowsPerWarp = 2;
constexpr int kTileRows = kScanWarps * kRowsPerWarp;          // 32
constexpr int kTileSmemBytes = kTileRows * kRowBytes;         // 49152
constexpr int kTileHalo = 64;                     // bytes staged after the tile for k-mer tails
constexpr int kTileMaxLen = kTileSmemBytes - 16;  // leaves room for the alignment skew
constexpr int kQueueCap = 544;                    // per-warp candidate queue (u16 entries)
constexpr int kRegionCache = 32;                  // regions of a tile kept in shared memory
constexpr uint32_t kNoOverflow = 0xFFFFFFFFu;
constexpr int kWindowSlots = 128;                 // per-tile window counter slots in shared memory
constexpr int kFlagTabBytes = 24576;              // direct stage-2 tables (sum 4^len, len<=7)

// ---- geometry of the lean kernel (scan_fast.cu): same tiles and rows, one warp per row -------------
constexpr int kFastThreads = 1024;
constexpr int kFastWarps = kFastThreads / 32;     // == kTileRows
constexpr int kFastQueueCap = 288;                // per-warp candidate queue (u16 entries), 6 spans
constexpr int kFastPackedWords = kRowBytes / 16 + 4;   // per-warp 2-bit copy of its row + halo word
constexpr int kFastFlagBytes = 16384;             // one direct table, pattern length <= 7
constexpr uint32_t kFastMinWindow = 512;          // <= 98 windows per tile, <= 1 window end per span

// ---- geometry of the window-row kernel (scan_rows.cu): a row is one window, a lane owns 32 bases ----
#ifndef TLS_ROW_THREADS
#define TLS_ROW_THREADS 768
#endif
#ifndef TLS_ROWS_PER_UNIT
#define TLS_ROWS_PER_UNIT 16
#endif
// measured on the 3.03 Gbp batch (threads / rows per unit: scan ms): 1024/8 1.18, 1024/16 1.20, 896/16 1.19, 768/8 1.17,
// 768/16 1.13, 768/32 1.13, 640/16 1.19, 512/16 1.28.  Other values: make EXTRA=-DTLS_ROW_THREADS=... OUT=... BUILD=...
constexpr int kRowThreads = TLS_ROW_THREADS;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kRowSpan = 32;                      // bases per lane
constexpr int kRowStage = 32 * kRowSpan;          // bytes of one staging buffer (1024)
constexpr int kRowMaxLen = kRowStage - 15 - 7;    // longest window: alignment skew <= 15, 7 bases of k-mer tail
constexpr uint32_t kRowMinWindow = 600;           // below this more than 40 % of the lanes would idle
constexpr uint32_t kRowMinWindows = 8;            // shorter contigs go to the general kernel
constexpr int kRowQueueCap = 256;                 // per-warp match queue (u16 entries), 8 lanes' worth
#ifndef TLS_ROW_KEYMODE
#define TLS_ROW_KEYMODE 1
#endif
constexpr int kRowKeyMode = TLS_ROW_KEYMODE;            // how a probe extracts its 16-bit key (probeKey in scan_rows.cu)
#ifndef TLS_ROW_PREFETCH
#define TLS_ROW_PREFETCH 1
#endif
constexpr bool kRowPrefetch = TLS_ROW_PREFETCH != 0;   // L2 prefetch of the row after next
constexpr int kRowsPerUnit = TLS_ROWS_PER_UNIT;   // consecutive windows one warp takes at a time (divides kTileRows)

// A maximal stretch of one item inside which matches must lie entirely
// (genome mode: the whole contig; tips mode: [0,T) and [n-T,n); reads: the whole read).
struct DevRegion {
    uint64_t bufStart;     // offset of the region's first base in the sequence buffer
    uint32_t len;
    uint32_t item;
    uint32_t relStart;     // position of the region's first base inside the item
    uint32_t itemLen;
    uint32_t winBase;      // global index of the item's first window (genome mode)
    uint32_t pad;
};

struct DevTile {
    uint64_t bufStart;     // first base of the tile
    uint32_t len;          // <= kTileMaxLen
    uint32_t firstRegion;  // first region that overlaps the tile
};

// Descriptor of a tile for the lean kernel, precomputed on the host: the tile lies inside one
// region (= one contig in genome mode).
struct DevFastTile {
    uint64_t bufStart;     // first base of the tile in the sequence buffer
    uint32_t len;          // <= kTileMaxLen
    uint32_t tile;         // index in the batch's tile table (slot addressing)
    uint32_t relBase;      // item-relative position of the tile's first base
    uint32_t regionEndQ;   // tile position one past the region's last base (>= len)
    uint32_t item, itemLen;
    uint32_t firstWin;     // global index of the window that holds the tile's first base
    uint32_t winDelta;     // item's first window - firstWin (mod 2^32): slot = rel / window + winDelta
    uint32_t termMixed;    // 1: matches of this tile may be "terminal" (src/teloscope.cpp:451-459)
    uint32_t pad;
};

// Up to kRowsPerUnit consecutive windows of one contig: the unit of work of the window-row kernel.
// Window k of the unit is global window firstWin + k, row slot slot0 + k, starts at item position
// rel + k * windowSize and at buffer offset bufStart + k * windowSize.
struct DevRowUnit {        // 32 bytes, read as two uint4
    uint64_t bufStart;
    uint32_t rel;          // multiple of windowSize
    uint32_t itemLen;
    uint32_t firstWin;
    uint32_t slot0;        // tile * kTileRows + row of the first window
    uint32_t nRows;        // 1..kRowsPerUnit
    uint32_t item;
};

// A contig the window-row kernel takes: its units and tiles are laid out on the device by
// rowPlanKernel (4 units per tile of 32 windows; the last tile of a contig may hold fewer).
struct DevRowRegion {      // 48 bytes
    uint64_t bufStart;
    uint32_t len, relStart, itemLen, winBase, item;
    uint32_t tile0;        // index of the contig's first tile in the batch's tile table
    uint32_t unit0;        // index of its first unit
    uint32_t nUnits;
    uint32_t pad[2];
};

struct DevMatch {          // 8 bytes, ordered by (item, pos, len)
    uint32_t pos;          // relative to the item start
    uint16_t len;
    uint8_t flags;         // bit0 forward, bit1 canonical, bit2 terminal (src/teloscope.cpp:451-459)
    uint8_t region;        // scan -> gather only: index of the match's region in its tile's region
                           // list, or kRegionAccounted when the scan kernel already did the counting
};
enum : uint8_t { kMatchFwd = 1, kMatchCan = 2, kMatchTerm = 4 };
constexpr uint8_t kRegionAccounted = 0xFF;

struct DevItemCounts {     // per item, filled by the scan kernel
    uint32_t all, fwd, rev, can, nonCanTerm;
    uint32_t records;      // window-row kernel only: match records of the item in the arena (sparse units emit canonical matches only)
    uint32_t pad[2];
};

struct DevBlock {          // 48 bytes
    uint32_t item;
    uint32_t order;        // terminal: bit31 = q-arm walk, low bits = emission order; ITS: 0
    uint32_t start;        // relative to the item start
    uint32_t len, counts, fwdCount, canCount, totalCov, fwdCov, canCov;
    uint8_t validOr;
    char label;
    uint16_t kind;         // 0 terminal, 1 interstitial
    uint32_t pad2;
};

struct DevLenClass {
    uint32_t len;
    uint32_t direct;       // 1: flag table in shared memory at flagOffset
    uint32_t flagOffset;
    uint32_t sortedBegin, sortedCount;
};

struct DevConfig {
    uint32_t nClasses, minLen, maxLen, npos;
    DevLenClass cls[8];
    uint32_t flagTabBytes;
    uint32_t mode;                 // tls_mode
    uint32_t windowSize, step, terminalLimit;
    uint32_t maxMatchDist, minBlockLen, maxBlockDist, minBlockCounts, itsMinLen;
    float minBlockDensity;
    uint32_t countAcgt;
    uint32_t overlapWindows;       // 1: step < windowSize.  The scan kernels then neither cut matches at window ends nor
                                   // account windows (every match of the contig is reported exactly once, see
                                   // overlapWindowsKernel in block_kernels.cu); the window table is built from the ordered list
    // rel / windowSize by multiply-high (Granlund-Montgomery, exact for every u32 rel):
    //   t = umulhi(winMagic, rel);  q = (t + ((rel - t) >> winSh1)) >> winSh2
    uint32_t winMagic, winSh1, winSh2;
};

struct ScanParams {
    const uint8_t *seq;            // device sequence buffer
    const DevTile *tiles;
    const DevRegion *regions;
    uint32_t nTiles, nRegions;
    const uint8_t *index8;         // 65536-byte stage-1 table
    const uint8_t *flagTabs;       // direct stage-2 tables
    const uint64_t *sortedKeys;    // long patterns (len > 7)
    const uint8_t *sortedFlags;
    // Match output: every (tile, warp) owns a slot of slotCap records; a warp that needs more
    // takes one overflow chunk of ovChunk records from ovArena (bump allocation).
    // work list: whole tiles (workTiles[nWork], or tiles 0..nWork-1 when null), then the rows the
    // lean kernel handed back (redoList[*redoCount] entries tile << 5 | row)
    const uint32_t *workTiles;
    uint32_t nWork, redoCap;
    const uint32_t *redoList;
    const unsigned int *redoCount;
    DevMatch *slots;               // [nTiles * kTileRows][slotCap]: one slot per tile row
    uint32_t slotCap;
    uint4 *slotMeta;               // [nTiles * kTileRows] x = records produced, y = overflow chunk base or 0xFFFFFFFF,
                                   // z / w = canonical / terminal non-canonical records among them
    uint32_t *tileTot;             // [3][nTiles] per-tile totals (all, canonical, terminal non-canonical), atomics
    DevMatch *ovArena;
    uint32_t ovCap, ovChunk;
    unsigned int *ovCursor;        // zeroed per launch; > ovCap afterwards means "grow and rerun"
    uint32_t *windows;             // [nWindows][8], zeroed per launch
    DevItemCounts *itemCounts;     // zeroed per launch
    DevConfig cfg;
};

struct FastParams {
    const uint8_t *seq;
    const DevFastTile *tiles;
    uint32_t nTiles;
    const uint8_t *index8;         // 65536-byte stage-1 table
    const uint8_t *flagTab;        // direct table of the single length class (4^len bytes)
    uint32_t flagTabBytes;         // rounded up to 16
    uint32_t patLen;
    DevMatch *slots;
    uint32_t slotCap;
    uint4 *slotMeta;               // per row slot: count, overflow base, canonical, terminal non-canonical
    uint32_t *tileTot;             // [3][nTilesAll]
    uint32_t nTilesAll;            // tiles of the batch (lean + general)
    DevMatch *ovArena;
    uint32_t ovCap, ovChunk;
    unsigned int *ovCursor;
    uint32_t *windows;             // [nWindows][8]
    DevItemCounts *itemCounts;
    uint32_t *redoList;            // rows with a non-ACGT byte: tile << 5 | row, redone by the general kernel
    unsigned int *redoCount;
    uint32_t redoCap;
    DevConfig cfg;
};

struct RowParams {
    const uint8_t *seq;
    const DevRowUnit *units;
    uint32_t nUnits;
    const uint8_t *indexRow;       // 65536 bytes: per 8-mer, NPOS-bit fields hit | forward << NPOS, and bit 2 NPOS = a canonical match among them
    uint32_t two16;                // 65536 (a multiplier the compiler cannot fold, see probeKey in scan_rows.cu)
    const uint8_t *flagTab;        // direct table of the single length class (4^len bytes)
    uint32_t flagTabBytes;         // rounded up to 16
    uint32_t patLen;
    DevMatch *slots;
    uint32_t slotCap;
    uint4 *slotMeta;               // per row slot: records, overflow base, canonical, terminal non-canonical (zeroed before the launch)
    uint32_t *tileTot;             // [3][nTilesAll]: records, canonical, terminal non-canonical
    uint32_t nTilesAll;
    DevMatch *ovArena;
    uint32_t ovCap, ovChunk;
    unsigned int *ovCursor;
    uint4 *windows;                // [nWindows][2]: A, C, G, T | canonical, non-canonical, forward, reverse covered bp
    uint4 *windows16;              // TLS_OUT_COMPACT: [nWindows] tls_window16 written directly (windows is not touched then)
    uint2 *windows8;               // TLS_OUT_PACKED: [nWindows] tls_window8 written directly;
    uint4 *winExc;                 // windows with a match count >= 255 (window, canonical, all, forward), appended at
    unsigned int *winExcCount;     // this counter
    uint32_t *zeroLanes;           // [nWindows]: bit l = lane l (row positions 32 l .. 32 l + 31) holds no match start
    DevItemCounts *itemCounts;
    unsigned int *status;          // set to 1 when the shared memory layout assumption fails
    unsigned int *unitCursor;      // work counter of this launch, zeroed before it
    unsigned int *partDone;        // progress: partDone[u / unitsPerPart] counts the finished units (the copy stream waits on these words)
    uint32_t unitsPerPart;         // 0: no progress counters
    DevConfig cfg;
};

// Per item, for the kernels behind the scan: where its sequence and its windows start, and whether the window-row kernel
// scanned it (then its record list is complete only outside sparseRange(), see device_common.cuh).
struct DevItemInfo {       // 16 bytes
    uint64_t bufStart;     // offset of the item's first base in the sequence buffer
    uint32_t winBase;      // global index of its first window (genome mode)
    uint32_t rowItem;      // 1: scanned by the window-row kernel
};

// ---- read scorer (scan_reads.cu): one warp per read, kReadStride start positions per row ----
constexpr int kReadThreads = 768;
constexpr int kReadStride = 1000;                 // <= 1024 - 15 (alignment skew) - 6 (bases a match needs beyond its start)

struct ReadParams {
    const uint8_t *seq;
    const DevItemInfo *items;      // bufStart per read
    const uint32_t *itemLen;       // after the '\r' trim (src/read-filter.cpp:38-40)
    uint32_t nItems;
    const uint8_t *index8;         // 65536-byte stage-1 table
    const uint8_t *flagTab;        // direct table of the single length class: bit0 hit, bit1 forward, bit2 canonical
    uint32_t flagTabBytes;         // rounded up to 16
    uint32_t patLen;
    uint8_t *pass;                 // per read: 1 = some terminal block (ReadTelomereFilter::matches)
    unsigned int *cursor;          // work counter, zeroed per launch
    unsigned int *status;          // set to 1 when the shared memory layout assumption fails
    unsigned long long *matchTotal;   // matches found (reads that are decided early stop counting)
    DevConfig cfg;
};

struct GatherParams {
    const DevTile *tiles;
    const DevRegion *regions;
    uint32_t nRegions;
    uint32_t *windows;             // [nWindows][8]: fields 4..7 (covered bp) of general-kernel records are added here
    DevItemCounts *itemCounts;
    DevConfig cfg;
    const DevMatch *slots;
    uint32_t slotCap;
    const uint4 *slotMeta;
    const DevMatch *ovArena;
    uint32_t ovCap;
    uint32_t nTiles;
    const unsigned long long *tileBase;   // [3][nTiles + 1] exclusive prefixes of the per-tile totals
    DevMatch *matches;             // ordered output
    unsigned long long matchCap;
    uint32_t *canIdx, *nctIdx;     // positions in `matches` of the canonical / terminal non-canonical records
    unsigned long long selCap;     // capacity of the canonical index list (and of sel4Can)
    unsigned long long selCapNct;  // capacity of the terminal non-canonical one (and of sel4Nct)
    uint32_t *sel4Can, *sel4Nct;   // TLS_OUT_PACKED: the final tls_match4 lists, written by the gather itself (else null)
    uint32_t *denseTiles;          // tiles with many records (dense units, other scan kernels): listed by the first gather kernel,
    unsigned int *denseCount;      // copied by the second one with a block per tile (capacity nTiles)
};

struct BlockParams {
    const DevMatch *matches;
    const unsigned long long *itemBegin;   // nItems + 1
    const DevItemCounts *itemCounts;
    const uint32_t *itemLen;               // per item (after the read-mode '\r' trim)
    uint32_t nItems;
    // window-row items: what a walk needs where the record list is sparse (null / 0 when the batch has none)
    const DevItemInfo *items;
    const uint8_t *seq;                    // the batch's sequence buffer
    const uint8_t *flagTab;                // direct table of the single length class: bit0 hit, bit1 forward, bit2 canonical
    uint32_t patLen;
    const uint32_t *zeroLanes;             // per window: lanes (32 positions each) without a match start
    DevBlock *blocks;                      // output, unordered; the host sorts by (item, order)
    uint32_t blockCap;
    unsigned int *blockCount;              // keeps counting past the capacity
    uint32_t *fwdBoundary, *revBoundary;   // per item, relative to the item start
    uint8_t *pass;                         // read mode: per item, 1 = some terminal block
    DevConfig cfg;
};

}  // namespace tls
