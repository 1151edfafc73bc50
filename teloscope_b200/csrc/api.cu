// C ABI of the B200-native scan (include/teloscope_b200.h): context, pinned double-buffered
// staging, tiling of the batch into scan regions, kernel sequencing, result assembly.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/teloscope_b200.h"
#include "kernels.hpp"
#include "patterns.hpp"

namespace tls {

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// TLS_TRACE=1 prints host-side phase timings of submit/collect to stderr (diagnostics only).
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(const char *what) : on(std::getenv("TLS_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {
        if (on) fprintf(stderr, "[tls] %s:", what);
    }
    void mark(const char *phase) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, " %s=%.2fms", phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
    ~Trace() { if (on) fprintf(stderr, "\n"); }
};

#define TLS_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(TLS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

// Growable device / pinned-host buffers -----------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool view = false;                       // p points into another buffer (the per-batch metadata block): nothing to free
    cudaError_t reserve(size_t bytes) {
        if (view) { p = nullptr; cap = 0; view = false; }
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void setView(void *q) {
        if (p && !view) cudaFree(p);
        p = q; cap = 0; view = true;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; cap = 0; view = false; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};
struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct Slot {
    bool busy = false, collected = false;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;       // window / item counters go home while the block kernels run
    cudaStream_t selStream = nullptr;        // -m lists: final records and their copy home, beside the block kernels
    cudaStream_t itsStream = nullptr;        // interstitial blocks of the items with complete lists, beside those of the window-row items
    cudaEvent_t evTerminal = nullptr, evItsLists = nullptr;
    cudaEvent_t evGather = nullptr;
    cudaEvent_t evTotals = nullptr;          // the tile scan has published the batch's totals (what tls_collect waits for first)
    uint32_t planLaunches = 0;               // kernels launched by tls_submit in front of enqueueScan (unit plan, tile scatter)
    bool specGather = false;                 // the gather was enqueued behind the scan with the buffers of the previous batch ...
    uint64_t specMatchCap = 0, specSelCap = 0, specNctCap = 0;   // ... of these capacities: kept if the totals fit, redone after growing them if not
    cudaEvent_t evScanDone = nullptr, evCopyDone = nullptr;
    static constexpr int kMaxParts = 8;      // launches of the window-row kernel on a large batch (result copy overlaps the scan)
    cudaEvent_t evPart[kMaxParts] = {};
    cudaStream_t partStream[kMaxParts] = {};  // launches 1 .. of the window-row kernel: side by side with launch 0, so that their CTAs fill
                                              // the SMs the previous launch's last units leave idle
    cudaEvent_t evSetup = nullptr;
    unsigned long long *hScanDev = nullptr;   // device view of hScan (mapped): scan totals without a copy
    cudaEvent_t evScan0 = nullptr, evScan1 = nullptr, evFast1 = nullptr, evAll0 = nullptr, evAll1 = nullptr;
    cudaEvent_t evDbg[6] = {};               // TLS_TRACE=1: timeline of the collect phase (gather, terminal, ITS, list copies, window copies)
    float fastMs = 0;
    // inputs
    uint32_t nItems = 0, mode = 0;
    uint64_t seqBytes = 0, nBases = 0, nWindows = 0;
    std::vector<uint64_t> absPos;
    std::vector<uint32_t> itemLen;
    std::vector<uint64_t> winOff;            // per item window offsets (n+1)
    std::vector<DevRegion> regions;
    std::vector<DevTile> tiles;
    std::vector<DevFastTile> fastTiles;      // tiles for the lean kernel
    std::vector<DevRowRegion> rowRegions;    // contigs of the window-row kernel (units are laid out on the device)
    uint32_t nRowUnits = 0;
    size_t nRowTiles = 0;
    std::vector<uint32_t> tileIndex;         // s.tiles holds the tiles of the lean / general kernels only; tileIndex[k] = index of
                                             // s.tiles[k] in the batch's tile table (the window-row kernel's tiles have no entry)
    uint32_t nTilesAll = 0;                  // tiles of the batch, window-row tiles included
    std::vector<uint32_t> generalTiles;      // indices of the tiles the general kernel takes entirely
    const uint8_t *devSeq = nullptr;         // sequence on the device (own buffer or caller's)
    HostBuf hSeq, hMeta, hMeta2;              // pinned staging
    DevBuf dRowTiles, dRowRegions, dFastTiles, dGeneralTiles, dRedo, dTileTot, dCanIdx, dNctIdx, dRunStarts;
    DevBuf dZeroLanes, dItemInfo, dConn, dWinRanges;      // window-row items: zero-lane masks per window, per-item info, ITS scratch
    std::vector<DevItemInfo> itemInfo;
    std::vector<uint2> winRanges;            // windows NOT written by the window-row kernel: [x, y) ranges ...
    std::vector<uint32_t> winRangePrefix;    // ... and the exclusive prefix of their lengths (nRanges + 1)
    uint32_t nRowItems = 0;
    HostBuf hMeta3;
    DevBuf dWinExc;                          // TLS_OUT_PACKED: windows whose match counts do not fit tls_window8
    HostBuf hWinExc;
    uint32_t nWinExc = 0;
    DevBuf dDenseTiles;                      // gather: tiles left to the block-per-tile kernel
    DevBuf dMetaAll;                         // one block per batch: every read-only table of the scan, uploaded with one copy
    bool fusedReads = false;                 // TLS_MODE_READS batch scored by the one-warp-per-read kernel (scan_reads.cu)
    uint64_t nCanAll = 0, nNctAll = 0;       // canonical / terminal non-canonical records of the batch
    DevBuf dAbsPos, dSeq, dRegions, dTiles, dItemLen, dItemCounts, dItemBegin, dWindows, dMatches, dSlots, dSlotMeta,
        dOvArena, dTileBase, dCounters, dBlocks, dBoundaries, dPass, dSelScratch, dSelCan, dSelNon, dWindows16;
    uint32_t ovCap = 0;
    // device counters block: [0] ovCursor(u64) [16] blockCount(u32) [20] ITS run count [24] ITS chain count [28] dense tiles of the gather [48] window exceptions (TLS_OUT_PACKED)
    HostBuf hScan, hCounters, hWindows, hItemCounts, hBlocks, hPass, hSelCan, hSelNon;
    uint64_t matchCap = 0;
    uint32_t blockCap = 0;
    uint32_t launches = 0;
    // assembled results
    std::vector<uint64_t> terminalOff, itsOff, canOff, nonOff;
    std::vector<tls_block> terminal, its;
    uint64_t nSelCan = 0, nSelNon = 0;
    std::vector<uint32_t> nCan, nFwd, nRev;
    std::vector<uint8_t> pass;
    uint64_t nMatches = 0;
    float scanMs = 0, allMs = 0;
};

}  // namespace
}  // namespace tls

using namespace tls;

struct tls_ctx {
    int device = 0;
    int smCount = 0;
    tls_config cfg{};
    std::vector<Pattern> patterns;
    ScanTables tables;
    DevConfig dcfg{};
    DevBuf dIndex, dFlags, dKeys, dKeyFlags;
    DevBuf dIndex16;                         // window-row kernel: one byte per 8-mer (hit | forward fields per start offset, canonical bit)
    uint32_t slotCap = 0, ovChunk = 0;
    std::string overlapRefusal;              // non-empty: genome-mode batches are refused (see buildDevConfig)
    bool forceGeneral = false;               // TLS_FORCE_GENERAL=1 or a lean-kernel layout failure
    bool noRowKernel = false;                // TLS_NO_ROW_KERNEL=1: 1536-base-row lean kernel instead of the window-row kernel
    bool noSpecGather = false;               // TLS_NO_SPEC_GATHER=1: the gather is always enqueued by tls_collect (A/B tests)
    bool noReadKernel = false;               // TLS_NO_READ_KERNEL=1: reads through the general scan + block kernels (A/B tests)
    // cuStreamWaitValue32 through the runtime's driver entry point lookup (no link against libcuda): the window copies of a
    // finished range of units are released by a stream wait on a progress word the scan kernel counts up, so ONE launch
    // scans the batch while its results already go home (TLS_ROW_PARTS=n launches instead when it is not available)
    int (*streamWaitValue32)(cudaStream_t, unsigned long long, unsigned int, unsigned int) = nullptr;
    int rowRanges = 8;                       // TLS_ROW_RANGES=n: ranges of units whose windows go home separately (1..8)
    bool zeroCopyWindows = false;            // TLS_OUT_COMPACT: the window-row kernel stores its 16-byte window records straight into the
                                             // pinned host buffer (posted writes over PCIe during the scan): measured slower (scan 1.23 ms against 0.95 + copy), off unless TLS_WINDOWS_ZEROCOPY=1
    bool compact = false;                    // TLS_OUT_COMPACT: tls_window16 / tls_match8 results
    bool packed = false;                     // TLS_OUT_PACKED: tls_window8 / tls_match4 results (compact is set too: one packed device buffer)
    int rowParts = 3;                        // TLS_ROW_PARTS=n (1..8): launches of the window-row kernel on a large batch
    uint32_t rowMinWindows = kRowMinWindows; // TLS_ROW_MIN_WINDOWS=n: shortest contig (in windows) the window-row kernel takes
    Slot slot[2];
    // pageable host input: a ring of small pinned buffers carries it to the device chunk by chunk (memcpy of chunk
    // k + 1 overlaps the DMA of chunk k); no pinned buffer of the batch's size is ever allocated
    static constexpr int kRing = 4;
    static constexpr size_t kRingBytes = 8u << 20;
    HostBuf ring[kRing];
    cudaEvent_t ringEv[kRing] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

void fillDivider(DevConfig &c, uint32_t d) {
    // Granlund-Montgomery round-up method for u32 / invariant d >= 1
    uint32_t L = 0;
    while ((1ull << L) < d) ++L;
    const uint64_t m = ((1ull << 32) * ((1ull << L) - d)) / d + 1;
    c.winMagic = static_cast<uint32_t>(m);
    c.winSh1 = L < 1 ? L : 1;
    c.winSh2 = L > 1 ? L - 1 : 0;
}

int buildDevConfig(tls_ctx *ctx) {
    const ScanTables &t = ctx->tables;
    DevConfig &c = ctx->dcfg;
    c = DevConfig{};
    c.nClasses = static_cast<uint32_t>(t.classes.size());
    c.minLen = t.minLen; c.maxLen = t.maxLen; c.npos = t.npos;
    for (size_t i = 0; i < t.classes.size(); ++i) {
        c.cls[i].len = t.classes[i].len;
        c.cls[i].direct = t.classes[i].direct ? 1u : 0u;
        c.cls[i].flagOffset = t.classes[i].flagOffset;
        c.cls[i].sortedBegin = t.classes[i].sortedBegin;
        c.cls[i].sortedCount = t.classes[i].sortedCount;
    }
    c.flagTabBytes = static_cast<uint32_t>((t.flagTables.size() + 15) & ~size_t(15));
    if (c.flagTabBytes > static_cast<uint32_t>(kFlagTabBytes))
        return fail(TLS_ERR_INVALID, "direct pattern tables need %u bytes (> %d)", c.flagTabBytes, kFlagTabBytes);
    const tls_config &u = ctx->cfg;
    c.windowSize = u.window_size; c.step = u.step; c.terminalLimit = u.terminal_limit;
    c.maxMatchDist = u.max_match_dist; c.minBlockLen = u.min_block_len; c.maxBlockDist = u.max_block_dist;
    c.minBlockCounts = u.min_block_counts; c.itsMinLen = t.itsMinLen; c.minBlockDensity = u.min_block_density;
    c.countAcgt = (u.outputs & TLS_OUT_ACGT) ? 1u : 0u;
    c.overlapWindows = u.step < u.window_size ? 1u : 0u;
    if (c.overlapWindows) {
        // src/teloscope.cpp:413-415 computes the scan start of a non-first window as min(step - L, overlap - L) in
        // uint32: with step < L or overlap < L it wraps and whole windows go unscanned; with several pattern lengths
        // the reference's match vectors are ordered by the window a match ENDS in, not by position.  Neither is
        // reproduced on the device; everything else (one pattern length, step >= L, window - step >= L) is.
        // Windows only exist in genome mode, so the refusal waits for a genome-mode batch (tls_submit).
        char why[160] = "";
        if (c.nClasses != 1)
            snprintf(why, sizeof why, "overlapping windows (step < window) need patterns of one length on the device");
        else if (u.step < c.maxLen || u.window_size - u.step < c.maxLen)
            snprintf(why, sizeof why, "overlapping windows need step >= %u and window - step >= %u (the pattern length)",
                     c.maxLen, c.maxLen);
        ctx->overlapRefusal = why;
    }
    ctx->packed = (u.outputs & TLS_OUT_PACKED) != 0;
    if (ctx->packed) {
        if (c.nClasses != 1 || u.window_size > 1023u)
            return fail(TLS_ERR_INVALID, "TLS_OUT_PACKED needs patterns of one length and window_size <= 1023 (10-bit count fields)");
    }
    ctx->compact = ctx->packed || (u.outputs & TLS_OUT_COMPACT) != 0;
    if (ctx->compact && !ctx->packed) {
        // a window's covered-bp counters are bounded by window_size x (sum of the pattern lengths): one match per length and start
        uint64_t lenSum = 0;
        for (uint32_t i = 0; i < c.nClasses; ++i) lenSum += c.cls[i].len;
        if (static_cast<uint64_t>(u.window_size) * lenSum > 65535ull)
            return fail(TLS_ERR_INVALID, "TLS_OUT_COMPACT needs window_size x (sum of pattern lengths) <= 65535 (16-bit window counters)");
    }
    fillDivider(c, u.window_size);
    return TLS_OK;
}

int upload(DevBuf &b, const void *src, size_t bytes) {
    TLS_CUDA(b.reserve(std::max<size_t>(bytes, 16)));
    if (bytes) TLS_CUDA(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
    return TLS_OK;
}

// Split the batch into scan regions and tiles (host, O(items + bases/48K)).
int planBatch(const tls_ctx *ctx, const tls_batch *b, Slot &s) {
    const uint32_t n = b->n_items;
    (void)ctx;
    s.regions.clear(); s.tiles.clear();
    s.itemLen.assign(n, 0);
    s.absPos.assign(n, 0);
    s.winOff.assign(n + 1, 0);
    s.nBases = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t a = b->offsets[i], e = b->offsets[i + 1];
        if (e < a) return fail(TLS_ERR_INVALID, "offsets not monotone at item %u", i);
        uint64_t len = e - a;
        if (len > 0xFFFFFFFFull) return fail(TLS_ERR_INVALID, "item %u longer than 2^32-1 bases", i);
        s.absPos[i] = b->abs_pos ? b->abs_pos[i] : 0;
        s.itemLen[i] = static_cast<uint32_t>(len);    // read mode '\r' trim happens in submit (host) below
    }
    return TLS_OK;
}

}  // namespace

extern "C" {

const char *tls_last_error(void) { return g_err.c_str(); }

static int tls_expand_patterns_impl(const char *raw_csv, int edit_distance, const char *canonical_fwd, char *text,
                        size_t text_len, uint8_t *is_forward, uint8_t *is_canonical, uint32_t cap,
                        uint32_t *n_out, size_t *text_need) {
    if (!raw_csv || !canonical_fwd || !n_out) return fail(TLS_ERR_INVALID, "null argument");
    std::vector<std::string> raw;
    std::string cur;
    for (const char *p = raw_csv;; ++p) {
        if (*p == ',' || *p == 0) {
            if (!cur.empty()) raw.push_back(cur);
            cur.clear();
            if (!*p) break;
        } else {
            cur.push_back(*p);
        }
    }
    std::vector<Pattern> pats;
    std::string err;
    if (!expandPatterns(raw, edit_distance, canonical_fwd, pats, err)) return fail(TLS_ERR_INVALID, "%s", err.c_str());
    size_t need = 0;
    for (const Pattern &p : pats) need += p.text.size() + 1;
    *n_out = static_cast<uint32_t>(pats.size());
    if (text_need) *text_need = need;
    if (!text) return TLS_OK;
    if (need > text_len || pats.size() > cap) return fail(TLS_ERR_INVALID, "output buffers too small");
    char *w = text;
    for (size_t i = 0; i < pats.size(); ++i) {
        std::memcpy(w, pats[i].text.c_str(), pats[i].text.size() + 1);
        w += pats[i].text.size() + 1;
        if (is_forward) is_forward[i] = pats[i].isForward;
        if (is_canonical) is_canonical[i] = pats[i].isCanonical;
    }
    return TLS_OK;
}

int tls_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp pr;
        if (cudaGetDeviceProperties(&pr, i) == cudaSuccess && pr.major == 10) ++ok;
    }
    return ok;
}

static int tls_create_impl(const tls_config *cfg, int device, tls_ctx **out) {
    if (!cfg || !out) return fail(TLS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != TLS_ABI_VERSION) return fail(TLS_ERR_INVALID, "ABI version mismatch");
    if (cfg->n_patterns == 0 || !cfg->patterns || !cfg->is_forward || !cfg->is_canonical)
        return fail(TLS_ERR_INVALID, "empty pattern table");
    if (cfg->window_size == 0 || cfg->step == 0 || cfg->step > cfg->window_size)
        return fail(TLS_ERR_INVALID, "need 0 < step <= window_size");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) {
        cudaGetLastError();
        return fail(TLS_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
    }
    if (device < 0 || device >= nDev) return fail(TLS_ERR_INVALID, "device %d out of range", device);
    cudaDeviceProp prop;
    TLS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TLS_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    TLS_CUDA(cudaSetDevice(device));

    // any failure below releases what was built so far (device buffers, streams, events) through tls_destroy
    struct Destroy { void operator()(tls_ctx *c) const { tls_destroy(c); } };
    std::unique_ptr<tls_ctx, Destroy> ctx(new (std::nothrow) tls_ctx);
    if (!ctx) return fail(TLS_ERR_NOMEM, "out of memory");
    ctx->device = device;
    ctx->smCount = prop.multiProcessorCount;
    ctx->cfg = *cfg;
    for (uint32_t i = 0; i < cfg->n_patterns; ++i) {
        Pattern p;
        p.text = cfg->patterns[i] ? cfg->patterns[i] : "";
        p.isForward = cfg->is_forward[i] != 0;
        p.isCanonical = cfg->is_canonical[i] != 0;
        ctx->patterns.push_back(std::move(p));
    }
    ctx->cfg.patterns = nullptr; ctx->cfg.is_forward = nullptr; ctx->cfg.is_canonical = nullptr;
    std::string err;
    if (!buildScanTables(ctx->patterns, ctx->tables, err)) return fail(TLS_ERR_INVALID, "%s", err.c_str());
    int rc = buildDevConfig(ctx.get());
    if (rc) return rc;
    const ScanTables &t = ctx->tables;
    if ((rc = upload(ctx->dIndex, t.index8.data(), t.index8.size()))) return rc;
    {
        std::vector<uint8_t> padded(ctx->dcfg.flagTabBytes, 0);
        std::copy(t.flagTables.begin(), t.flagTables.end(), padded.begin());
        if ((rc = upload(ctx->dFlags, padded.data(), padded.size()))) return rc;
    }
    if (t.classes.size() == 1 && t.classes[0].direct && t.classes[0].len <= 7) {
        // the window-row kernel's table: per 8-mer and start offset o < NP the hit and forward flags of the L-mer that starts
        // there, and one bit that says whether one of them is the canonical pattern
        const uint32_t L = t.classes[0].len, NP = rowNpos(L), mask = (1u << (2 * L)) - 1u;
        std::vector<uint8_t> idx(65536, 0);
        for (uint32_t key = 0; key < 65536; ++key) {
            uint32_t e = 0;
            for (uint32_t o = 0; o < NP; ++o) {
                const uint8_t f = t.flagTables[t.classes[0].flagOffset + ((key >> (2 * o)) & mask)];
                if (f & 1u) e |= (1u << o) | ((f & 2u) ? 1u << (NP + o) : 0u) | ((f & 4u) ? 1u << (2 * NP) : 0u);
            }
            idx[key] = static_cast<uint8_t>(e);
        }
        if ((rc = upload(ctx->dIndex16, idx.data(), idx.size()))) return rc;
    }
    if ((rc = upload(ctx->dKeys, t.sortedKeys.data(), t.sortedKeys.size() * 8))) return rc;
    if ((rc = upload(ctx->dKeyFlags, t.sortedFlags.data(), t.sortedFlags.size()))) return rc;
    {
        // expected matches per (tile, warp) on random sequence = bases per warp * pattern density;
        // slots hold twice that (power of two), denser stretches spill into overflow chunks
        double density = 0;
        for (const LenClass &c : t.classes) density += static_cast<double>(c.sortedCount) / static_cast<double>(1ull << (2 * std::min<uint32_t>(c.len, 31)));
        const uint32_t perRow = static_cast<uint32_t>(kRowBytes);          // every tile row owns one slot
        ctx->ovChunk = perRow * ctx->dcfg.nClasses;
        uint32_t want = static_cast<uint32_t>(1.5 * perRow * density) + 48, cap = 64;
        while (cap < want) cap <<= 1;
        ctx->slotCap = std::min(cap, ctx->ovChunk);
    }
    for (Slot &s : ctx->slot) {
        TLS_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        TLS_CUDA(cudaStreamCreateWithFlags(&s.copyStream, cudaStreamNonBlocking));
        TLS_CUDA(cudaStreamCreateWithFlags(&s.selStream, cudaStreamNonBlocking));
        TLS_CUDA(cudaStreamCreateWithFlags(&s.itsStream, cudaStreamNonBlocking));
        TLS_CUDA(cudaEventCreateWithFlags(&s.evTerminal, cudaEventDisableTiming));
        TLS_CUDA(cudaEventCreateWithFlags(&s.evItsLists, cudaEventDisableTiming));
        TLS_CUDA(cudaEventCreate(&s.evGather));
        TLS_CUDA(cudaEventCreateWithFlags(&s.evTotals, cudaEventDisableTiming));
        TLS_CUDA(cudaEventCreateWithFlags(&s.evScanDone, cudaEventDisableTiming));
        TLS_CUDA(cudaEventCreateWithFlags(&s.evCopyDone, cudaEventDisableTiming));
        TLS_CUDA(cudaEventCreate(&s.evScan0)); TLS_CUDA(cudaEventCreate(&s.evScan1)); TLS_CUDA(cudaEventCreate(&s.evFast1));
        TLS_CUDA(cudaEventCreate(&s.evAll0)); TLS_CUDA(cudaEventCreate(&s.evAll1));
        TLS_CUDA(s.dCounters.reserve(128));             // [64..96): work counters of the window-row launches
        TLS_CUDA(s.hCounters.reserve(64));
        TLS_CUDA(cudaHostAlloc(&s.hScan.p, 4096, cudaHostAllocMapped));
        s.hScan.cap = 4096;
        TLS_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s.hScanDev), s.hScan.p, 0));
    }
    {
        bool layoutOk = false;
        TLS_CUDA(probeRowLayout(ctx->slot[0].dCounters.as<unsigned int>(), layoutOk, ctx->slot[0].stream));
        if (!layoutOk) { ctx->noRowKernel = true; ctx->noReadKernel = true; }     // tile kernels take over: same results, slower
    }
    ctx->forceGeneral = std::getenv("TLS_FORCE_GENERAL") != nullptr;
    ctx->noRowKernel = ctx->noRowKernel || std::getenv("TLS_NO_ROW_KERNEL") != nullptr;
    ctx->noReadKernel = ctx->noReadKernel || std::getenv("TLS_NO_READ_KERNEL") != nullptr;
    ctx->noSpecGather = std::getenv("TLS_NO_SPEC_GATHER") != nullptr;
    if (!std::getenv("TLS_ROW_PARTS")) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess && fn)
            ctx->streamWaitValue32 = reinterpret_cast<int (*)(cudaStream_t, unsigned long long, unsigned int, unsigned int)>(fn);
        cudaGetLastError();
    }
    if (const char *e = std::getenv("TLS_ROW_RANGES")) ctx->rowRanges = std::min(static_cast<int>(Slot::kMaxParts), std::max(1, std::atoi(e)));
    ctx->zeroCopyWindows = std::getenv("TLS_WINDOWS_ZEROCOPY") != nullptr;
    if (const char *e = std::getenv("TLS_ROW_PARTS")) ctx->rowParts = std::min(static_cast<int>(Slot::kMaxParts), std::max(1, std::atoi(e)));
    if (const char *e = std::getenv("TLS_ROW_MIN_WINDOWS")) ctx->rowMinWindows = std::max(1, std::atoi(e));
    *out = ctx.release();
    return TLS_OK;
}

void tls_destroy(tls_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (Slot &s : ctx->slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        for (DevBuf *b : {&s.dRowTiles, &s.dRowRegions, &s.dFastTiles, &s.dGeneralTiles, &s.dRedo, &s.dTileTot, &s.dCanIdx,
                          &s.dNctIdx, &s.dRunStarts, &s.dZeroLanes, &s.dItemInfo, &s.dConn, &s.dWinRanges, &s.dMetaAll, &s.dDenseTiles, &s.dWinExc})
            b->release();
        for (DevBuf *b : {&s.dAbsPos, &s.dSeq, &s.dRegions, &s.dTiles, &s.dItemLen, &s.dItemCounts, &s.dItemBegin, &s.dWindows,
                          &s.dMatches, &s.dSlots, &s.dSlotMeta, &s.dOvArena, &s.dTileBase, &s.dCounters,
                          &s.dBlocks, &s.dBoundaries, &s.dPass, &s.dSelScratch, &s.dSelCan, &s.dSelNon, &s.dWindows16})
            b->release();
        for (HostBuf *b : {&s.hSeq, &s.hMeta, &s.hMeta2, &s.hMeta3, &s.hScan, &s.hCounters, &s.hWindows, &s.hWinExc, &s.hItemCounts, &s.hBlocks, &s.hPass,
                           &s.hSelCan, &s.hSelNon})
            b->release();
        if (s.evScan0) cudaEventDestroy(s.evScan0);
        if (s.evScan1) cudaEventDestroy(s.evScan1);
        for (cudaEvent_t &e : s.evDbg) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (s.evFast1) cudaEventDestroy(s.evFast1);
        if (s.evAll0) cudaEventDestroy(s.evAll0);
        if (s.evAll1) cudaEventDestroy(s.evAll1);
        if (s.copyStream) { cudaStreamSynchronize(s.copyStream); cudaStreamDestroy(s.copyStream); }
        if (s.selStream) { cudaStreamSynchronize(s.selStream); cudaStreamDestroy(s.selStream); }
        if (s.itsStream) { cudaStreamSynchronize(s.itsStream); cudaStreamDestroy(s.itsStream); }
        if (s.evTerminal) cudaEventDestroy(s.evTerminal);
        if (s.evItsLists) cudaEventDestroy(s.evItsLists);
        if (s.evGather) cudaEventDestroy(s.evGather);
        if (s.evTotals) cudaEventDestroy(s.evTotals);
        if (s.evScanDone) cudaEventDestroy(s.evScanDone);
        for (cudaEvent_t &e : s.evPart) if (e) { cudaEventDestroy(e); e = nullptr; }
        for (cudaStream_t &st : s.partStream) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); st = nullptr; }
        if (s.evSetup) cudaEventDestroy(s.evSetup);
        if (s.evCopyDone) cudaEventDestroy(s.evCopyDone);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (DevBuf *b : {&ctx->dIndex, &ctx->dIndex16, &ctx->dFlags, &ctx->dKeys, &ctx->dKeyFlags}) b->release();
    for (int k = 0; k < tls_ctx::kRing; ++k) {
        if (ctx->ringEv[k]) cudaEventDestroy(ctx->ringEv[k]);
        ctx->ring[k].release();
    }
    delete ctx;
}

}  // extern "C"

namespace {

GatherParams gatherParams(const tls_ctx *ctx, const Slot &s, uint64_t matchCap, uint64_t selCap, uint64_t selCapNct) {
    GatherParams g{};
    g.tiles = s.dTiles.as<DevTile>(); g.regions = s.dRegions.as<DevRegion>();
    g.nRegions = static_cast<uint32_t>(s.regions.size());
    g.windows = s.dWindows.as<uint32_t>(); g.itemCounts = s.dItemCounts.as<DevItemCounts>();
    g.cfg = ctx->dcfg; g.cfg.mode = s.mode;
    g.slots = s.dSlots.as<DevMatch>(); g.slotCap = ctx->slotCap;
    g.slotMeta = s.dSlotMeta.as<uint4>();
    g.ovArena = s.dOvArena.as<DevMatch>(); g.ovCap = s.ovCap;
    g.nTiles = s.nTilesAll;
    g.tileBase = s.dTileBase.as<unsigned long long>();
    g.matches = s.dMatches.as<DevMatch>(); g.matchCap = matchCap;
    g.canIdx = s.dCanIdx.as<uint32_t>(); g.nctIdx = s.dNctIdx.as<uint32_t>();
    g.selCap = selCap; g.selCapNct = selCapNct;
    const bool fusedLists = ctx->packed && s.mode == TLS_MODE_GENOME && (ctx->cfg.outputs & TLS_OUT_MATCHES);   // the buffers are reserved by then
    g.sel4Can = fusedLists ? s.dSelCan.as<uint32_t>() : nullptr;
    g.sel4Nct = fusedLists ? s.dSelNon.as<uint32_t>() : nullptr;
    g.denseTiles = s.dDenseTiles.as<uint32_t>();
    g.denseCount = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 28);
    return g;
}

// Slots -> ordered list (+ packed -m lists) and the per-item offsets.  The two gather kernels write disjoint places: side by
// side (the dense tiles were listed by the tile scan).
int enqueueGather(tls_ctx *ctx, Slot &s, uint64_t matchCap, uint64_t selCap, uint64_t selCapNct) {
    if (s.nTilesAll) {
        TLS_CUDA(cudaEventRecord(s.evTerminal, s.stream));
        TLS_CUDA(cudaStreamWaitEvent(s.itsStream, s.evTerminal, 0));
        TLS_CUDA(launchGather(gatherParams(ctx, s, matchCap, selCap, selCapNct), s.stream, s.itsStream));
        TLS_CUDA(cudaEventRecord(s.evItsLists, s.itsStream));
        TLS_CUDA(cudaStreamWaitEvent(s.stream, s.evItsLists, 0));
        s.launches += 2;
    }
    if (s.nItems) {
        TLS_CUDA(launchItemOffsets(s.dItemCounts.as<DevItemCounts>(), s.nRowItems ? s.dItemInfo.as<DevItemInfo>() : nullptr, s.nItems,
                                   s.dItemBegin.as<unsigned long long>(), s.stream));
        s.launches++;
    }
    TLS_CUDA(cudaEventRecord(s.evGather, s.stream));
    return TLS_OK;
}

// Enqueue the scan for the slot: memsets, lean kernel over the single-contig tiles, general kernel
// over the rest and over the rows the lean kernel handed back, per-tile prefix, copy of the counters.
int enqueueScan(tls_ctx *ctx, Slot &s) {
    static const bool traceOn = std::getenv("TLS_TRACE") != nullptr;
    const auto tEnter = std::chrono::steady_clock::now();
    const uint32_t nTiles = s.nTilesAll;
    const size_t nSlots = static_cast<size_t>(std::max<uint32_t>(nTiles, 1)) * kTileRows;
    TLS_CUDA(s.dSlots.reserve(nSlots * ctx->slotCap * sizeof(DevMatch)));
    TLS_CUDA(s.dSlotMeta.reserve(nSlots * 16));
    TLS_CUDA(s.dTileTot.reserve(static_cast<size_t>(std::max<uint32_t>(nTiles, 1)) * 12));
    TLS_CUDA(s.dTileBase.reserve((static_cast<size_t>(nTiles) + 1) * 24));
    TLS_CUDA(cudaMemsetAsync(s.dTileTot.p, 0, static_cast<size_t>(std::max<uint32_t>(nTiles, 1)) * 12, s.stream));
    if (s.nRowUnits) {
        // the last tile of a contig has fewer than 32 windows: its other row slots must read as empty
        TLS_CUDA(cudaMemsetAsync(s.dSlotMeta.p, 0, nSlots * 16, s.stream));
    }
    s.ovCap = std::max<uint32_t>(s.ovCap, 256u * ctx->ovChunk);
    TLS_CUDA(s.dOvArena.reserve(static_cast<size_t>(s.ovCap) * sizeof(DevMatch)));
    const uint32_t nFast = static_cast<uint32_t>(s.fastTiles.size());
    const uint32_t redoCap = std::max<uint32_t>(nFast, 1) * kTileRows;
    TLS_CUDA(s.dRedo.reserve(16 + static_cast<size_t>(redoCap) * 4));       // [0] count, [1] status, [4..] entries
    TLS_CUDA(cudaStreamWaitEvent(s.stream, s.evCopyDone, 0));               // a rerun must not clear what is still being copied
    TLS_CUDA(cudaMemsetAsync(s.dRedo.p, 0, 16, s.stream));
    TLS_CUDA(cudaMemsetAsync(s.dCounters.p, 0, 128, s.stream));
    TLS_CUDA(cudaMemsetAsync(s.dItemCounts.p, 0, std::max<size_t>(s.nItems, 1) * sizeof(DevItemCounts), s.stream));
    s.launches = s.planLaunches;                            // the plan kernels tls_submit launched for this batch
    // window counters: the window-row kernel writes final records; the other scan kernels add into zeroed 32-byte counters
    const uint2 *dRanges = s.dWinRanges.as<uint2>();
    const uint32_t *dRangePrefix = reinterpret_cast<const uint32_t *>(s.dWinRanges.as<uint8_t>() + s.winRanges.size() * sizeof(uint2));
    const uint32_t nRanges = static_cast<uint32_t>(s.winRanges.size()), nRangeWindows = s.winRangePrefix.back();
    if (s.nWindows && !s.nRowUnits) TLS_CUDA(cudaMemsetAsync(s.dWindows.p, 0, s.nWindows * 32, s.stream));
    else if (nRangeWindows) {
        TLS_CUDA(launchZeroWindowRanges(s.dWindows.as<uint32_t>(), dRanges, dRangePrefix, nRanges, nRangeWindows, s.stream));
        s.launches++;
    }
    if (nTiles) {
        DevConfig cfg = ctx->dcfg;
        cfg.mode = s.mode;
        if (s.mode != TLS_MODE_GENOME || cfg.overlapWindows) cfg.countAcgt = 0;   // overlapping windows: counted by overlapWindowsKernel
        unsigned int *redo = s.dRedo.as<unsigned int>();
        TLS_CUDA(cudaEventRecord(s.evScan0, s.stream));
        if (nFast) {
            FastParams f{};
            f.seq = s.devSeq;
            f.tiles = s.dFastTiles.as<DevFastTile>(); f.nTiles = nFast;
            f.index8 = ctx->dIndex.as<uint8_t>();
            f.flagTab = ctx->dFlags.as<uint8_t>() + cfg.cls[0].flagOffset;
            f.flagTabBytes = ((1u << (2 * cfg.cls[0].len)) + 15u) & ~15u;
            f.patLen = cfg.cls[0].len;
            f.slots = s.dSlots.as<DevMatch>(); f.slotCap = ctx->slotCap;
            f.slotMeta = s.dSlotMeta.as<uint4>();
            f.tileTot = s.dTileTot.as<uint32_t>(); f.nTilesAll = nTiles;
            f.ovArena = s.dOvArena.as<DevMatch>(); f.ovCap = s.ovCap; f.ovChunk = ctx->ovChunk;
            f.ovCursor = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>());
            f.windows = s.dWindows.as<uint32_t>();
            f.itemCounts = s.dItemCounts.as<DevItemCounts>();
            f.redoList = redo + 4; f.redoCount = redo; f.redoCap = redoCap;
            f.cfg = cfg;
            TLS_CUDA(launchFastScan(f, std::min<int>(ctx->smCount, static_cast<int>(nFast)), s.stream));
            s.launches++;
        }
        const uint32_t nGeneral = static_cast<uint32_t>(s.generalTiles.size());
        if (nGeneral || nFast) {
            ScanParams p{};
            p.seq = s.devSeq;
            p.tiles = s.dTiles.as<DevTile>();
            p.regions = s.dRegions.as<DevRegion>();
            p.nTiles = nTiles;
            p.nRegions = static_cast<uint32_t>(s.regions.size());
            p.workTiles = s.dGeneralTiles.as<uint32_t>(); p.nWork = nGeneral;
            p.redoList = redo + 4; p.redoCount = nFast ? redo : nullptr; p.redoCap = redoCap;
            p.index8 = ctx->dIndex.as<uint8_t>();
            p.flagTabs = ctx->dFlags.as<uint8_t>();
            p.sortedKeys = ctx->dKeys.as<uint64_t>();
            p.sortedFlags = ctx->dKeyFlags.as<uint8_t>();
            p.slots = s.dSlots.as<DevMatch>(); p.slotCap = ctx->slotCap;
            p.slotMeta = s.dSlotMeta.as<uint4>();
            p.tileTot = s.dTileTot.as<uint32_t>();
            p.ovArena = s.dOvArena.as<DevMatch>(); p.ovCap = s.ovCap; p.ovChunk = ctx->ovChunk;
            p.ovCursor = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>());
            p.windows = s.dWindows.as<uint32_t>();
            p.itemCounts = s.dItemCounts.as<DevItemCounts>();
            p.cfg = cfg;
            // rows handed back by the lean kernel are only known on the device: when there are lean
            // tiles every SM gets a CTA, which exits at once if the work list turns out empty
            const int grid = nFast ? ctx->smCount : std::min<int>(ctx->smCount, static_cast<int>(nGeneral));
            TLS_CUDA(launchScan(p, grid, s.stream));
            s.launches++;
        }
        // The window-row kernel goes last and, on a large batch, in ctx->rowParts launches over consecutive unit ranges: the
        // window counters of a finished range (and of every small contig before its end: the lean / general kernels
        // have already run) go home on the side stream while the next range is scanned.  The result copy of a
        // device-resident batch is PCIe-bound (32 bytes per window), so hiding part of it behind the scan shortens the
        // step; every extra launch costs the scan 30-40 us (table reload, ramp-up, tail): three by default (measured on the
        // 3.03 Gbp batch, parts / scan ms / step ms: 1 / 1.19 / 4.03, 2 / 1.27 / 3.45, 3 / 1.28 / 3.27, 4 / 1.31 / 3.28).
        // (One launch with progress counters and a polling kernel on the side stream was tried: it disturbed the scan.)
        const uint32_t nRow = s.nRowUnits;
        const bool anyWindows = s.mode == TLS_MODE_GENOME && s.nWindows && !cfg.overlapWindows;
        // what goes home per window: the 32-byte counters, or (TLS_OUT_COMPACT) 16-byte records: written directly by the
        // window-row kernel, packed from the counters for the windows of the other kernels
        const size_t winRec = ctx->packed ? 8 : (ctx->compact ? 16 : 32);
        if (anyWindows) {
            TLS_CUDA(s.hWindows.reserve(s.nWindows * winRec));
            if (ctx->compact) TLS_CUDA(s.dWindows16.reserve(s.nWindows * 16));
        }
        if (ctx->packed) TLS_CUDA(s.dWinExc.reserve(std::max<uint64_t>(s.nWindows, 1) * sizeof(tls_window_exc)));   // every window could be one
        unsigned int *winExcCount = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 48);
        const uint8_t *winSrc = ctx->compact ? s.dWindows16.as<uint8_t>() : s.dWindows.as<uint8_t>();
        // Compact records of a batch with window-row contigs: the kernels store them straight into the pinned host buffer
        // (48 MB per 3 Gbp trickle over PCIe while the scan runs, nothing is left to copy afterwards).
        const bool zeroCopy = anyWindows && ctx->compact && nRow && ctx->zeroCopyWindows;
        void *win16Dst = s.dWindows16.p;
        if (zeroCopy) TLS_CUDA(cudaHostGetDevicePointer(&win16Dst, s.hWindows.p, 0));
        if (anyWindows && ctx->compact && nRow && nRangeWindows) {
            if (ctx->packed) TLS_CUDA(launchPackWindow8Ranges(s.dWindows.as<uint32_t>(), win16Dst, dRanges, dRangePrefix, nRanges, nRangeWindows, cfg.cls[0].len, s.dWinExc.p, winExcCount, s.stream));
            else TLS_CUDA(launchPackWindowRanges(s.dWindows.as<uint32_t>(), win16Dst, dRanges, dRangePrefix, nRanges, nRangeWindows, s.stream));
            s.launches++;
        }
        uint64_t winCopied = 0;
        if (nRow) {
            RowParams f{};
            f.seq = s.devSeq;
            f.indexRow = ctx->dIndex16.as<uint8_t>();
            f.two16 = 65536u;
            f.windows16 = (ctx->compact && !ctx->packed) ? static_cast<uint4 *>(win16Dst) : nullptr;
            f.windows8 = ctx->packed ? static_cast<uint2 *>(win16Dst) : nullptr;
            f.winExc = s.dWinExc.as<uint4>(); f.winExcCount = winExcCount;
            f.zeroLanes = s.dZeroLanes.as<uint32_t>();
            f.flagTab = ctx->dFlags.as<uint8_t>() + cfg.cls[0].flagOffset;
            f.flagTabBytes = ((1u << (2 * cfg.cls[0].len)) + 15u) & ~15u;
            f.patLen = cfg.cls[0].len;
            f.slots = s.dSlots.as<DevMatch>(); f.slotCap = ctx->slotCap;
            f.slotMeta = s.dSlotMeta.as<uint4>();
            f.tileTot = s.dTileTot.as<uint32_t>(); f.nTilesAll = nTiles;
            f.ovArena = s.dOvArena.as<DevMatch>(); f.ovCap = s.ovCap; f.ovChunk = ctx->ovChunk;
            f.ovCursor = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>());
            f.windows = s.dWindows.as<uint4>();
            f.itemCounts = s.dItemCounts.as<DevItemCounts>();
            f.status = redo + 1;
            f.cfg = cfg;
            const uint32_t gridWarps = static_cast<uint32_t>(ctx->smCount) * kRowWarps;
            auto firstWindowOfUnit = [&](uint32_t u) {          // rowRegions are sorted by unit0
                size_t lo = 0, hi = s.rowRegions.size();
                while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (s.rowRegions[mid].unit0 <= u) lo = mid; else hi = mid; }
                return static_cast<uint64_t>(s.rowRegions[lo].winBase) + static_cast<uint64_t>(u - s.rowRegions[lo].unit0) * kRowsPerUnit;
            };
            const bool bigBatch = anyWindows && !zeroCopy && nRow >= 16u * gridWarps;
            const int ranges = (bigBatch && ctx->streamWaitValue32) ? ctx->rowRanges : 1;
            const int parts = (bigBatch && !ctx->streamWaitValue32) ? ctx->rowParts : 1;
            if (ranges > 1) {
                // ONE launch; the copy stream waits (stream wait-value, no host thread, no SM) for the progress word of each range
                // of units and sends that range's windows home while the kernel is still scanning the next ones
                const uint32_t upp = (nRow + static_cast<uint32_t>(ranges) - 1) / static_cast<uint32_t>(ranges);
                f.units = s.dRowTiles.as<DevRowUnit>(); f.nUnits = nRow;
                f.unitCursor = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 64);
                f.partDone = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 96);
                f.unitsPerPart = upp;
                if (!s.evSetup) TLS_CUDA(cudaEventCreateWithFlags(&s.evSetup, cudaEventDisableTiming));
                TLS_CUDA(cudaEventRecord(s.evSetup, s.stream));          // counters zeroed, small contigs' windows final
                TLS_CUDA(cudaStreamWaitEvent(s.copyStream, s.evSetup, 0));
                TLS_CUDA(launchRowScan(f, static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(ctx->smCount), (nRow + kRowWarps - 1) / kRowWarps)), s.stream));
                if (traceOn) fprintf(stderr, " [row kernel enqueued %.3f ms into enqueueScan]", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tEnter).count());
                s.launches++;
                for (int k = 0; k + 1 < ranges; ++k) {
                    const uint32_t uHi = std::min<uint32_t>(nRow, static_cast<uint32_t>(k + 1) * upp), uLo = static_cast<uint32_t>(k) * upp;
                    if (uLo >= nRow) break;
                    const int rc = ctx->streamWaitValue32(s.copyStream, reinterpret_cast<unsigned long long>(f.partDone + k), uHi - uLo, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
                    if (rc != 0) return fail(TLS_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", rc);
                    const uint64_t wEnd = uHi >= nRow ? s.nWindows : firstWindowOfUnit(uHi);
                    if (wEnd > winCopied)
                        TLS_CUDA(cudaMemcpyAsync(s.hWindows.as<uint8_t>() + winCopied * winRec, winSrc + winCopied * winRec,
                                                 (wEnd - winCopied) * winRec, cudaMemcpyDeviceToHost, s.copyStream));
                    winCopied = std::max(winCopied, wEnd);
                }
            } else {
            bool partsPending = false;
            for (int part = 0; part < parts; ++part) {
                const uint32_t uLo = static_cast<uint32_t>(static_cast<uint64_t>(nRow) * part / parts);
                const uint32_t uHi = static_cast<uint32_t>(static_cast<uint64_t>(nRow) * (part + 1) / parts);
                f.units = s.dRowTiles.as<DevRowUnit>() + uLo; f.nUnits = uHi - uLo;
                f.unitCursor = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 64) + part;
                const int grid = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(ctx->smCount), (f.nUnits + kRowWarps - 1) / kRowWarps));
                cudaStream_t st = s.stream;
                if (part > 0) {                               // its own stream, behind everything enqueued before the first launch
                    if (!s.partStream[part]) TLS_CUDA(cudaStreamCreateWithFlags(&s.partStream[part], cudaStreamNonBlocking));
                    st = s.partStream[part];
                    TLS_CUDA(cudaStreamWaitEvent(st, s.evSetup, 0));
                }
                if (part == 0 && parts > 1) {                 // the point the other launches wait for: just before launch 0
                    if (!s.evSetup) TLS_CUDA(cudaEventCreateWithFlags(&s.evSetup, cudaEventDisableTiming));
                    TLS_CUDA(cudaEventRecord(s.evSetup, s.stream));
                }
                TLS_CUDA(launchRowScan(f, grid, st));
                s.launches++;
                if (!s.evPart[part]) TLS_CUDA(cudaEventCreateWithFlags(&s.evPart[part], cudaEventDisableTiming));
                TLS_CUDA(cudaEventRecord(s.evPart[part], st));
                if (part > 0) partsPending = true;
                if (part + 1 < parts) {
                    const uint64_t wEnd = firstWindowOfUnit(uHi);   // every window below it is final once launches 0 .. part are done
                    TLS_CUDA(cudaStreamWaitEvent(s.copyStream, s.evPart[part], 0));
                    if (wEnd > winCopied)
                        TLS_CUDA(cudaMemcpyAsync(s.hWindows.as<uint8_t>() + winCopied * winRec, winSrc + winCopied * winRec,
                                                 (wEnd - winCopied) * winRec, cudaMemcpyDeviceToHost, s.copyStream));
                    winCopied = std::max(winCopied, wEnd);
                }
            }
            if (partsPending)                                 // the main stream goes on when every launch is done
                for (int part = 1; part < parts; ++part) TLS_CUDA(cudaStreamWaitEvent(s.stream, s.evPart[part], 0));
            }
        }
        TLS_CUDA(cudaEventRecord(s.evFast1, s.stream));
        TLS_CUDA(cudaEventRecord(s.evScan1, s.stream));
        if (anyWindows && ctx->compact && !nRow) {
            if (ctx->packed) TLS_CUDA(launchPackWindows8(s.dWindows.as<uint32_t>(), s.dWindows16.p, s.nWindows, cfg.cls[0].len, s.dWinExc.p, winExcCount, s.stream));
            else TLS_CUDA(launchPackWindows(s.dWindows.as<uint32_t>(), s.dWindows16.p, s.nWindows, s.stream));
            s.launches++;
        }
        // window and per-item counters are final now: copy (the rest of) them home on the side stream
        TLS_CUDA(cudaEventRecord(s.evScanDone, s.stream));
        TLS_CUDA(cudaStreamWaitEvent(s.copyStream, s.evScanDone, 0));
        if (anyWindows && !zeroCopy && winCopied < s.nWindows)
            TLS_CUDA(cudaMemcpyAsync(s.hWindows.as<uint8_t>() + winCopied * winRec, winSrc + winCopied * winRec,
                                     (s.nWindows - winCopied) * winRec, cudaMemcpyDeviceToHost, s.copyStream));
        if (s.mode != TLS_MODE_READS && s.nItems) {
            TLS_CUDA(s.hItemCounts.reserve(static_cast<size_t>(s.nItems) * sizeof(DevItemCounts)));
            TLS_CUDA(cudaMemcpyAsync(s.hItemCounts.p, s.dItemCounts.p, static_cast<size_t>(s.nItems) * sizeof(DevItemCounts),
                                     cudaMemcpyDeviceToHost, s.copyStream));
        }
        TLS_CUDA(cudaEventRecord(s.evCopyDone, s.copyStream));
        TLS_CUDA(s.dSelScratch.reserve(tileScanScratchBytes(nTiles)));
        TLS_CUDA(s.dDenseTiles.reserve(static_cast<size_t>(nTiles) * 4));
        TLS_CUDA(launchTileScan(s.dTileTot.as<uint32_t>(), nTiles, s.dTileBase.as<unsigned long long>(), s.dSelScratch.as<unsigned long long>(),
                                reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>()), (nFast || nRow) ? redo : nullptr,
                                s.hScanDev, s.dDenseTiles.as<uint32_t>(), reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 28), s.stream));
        s.launches += 2;
        TLS_CUDA(cudaEventRecord(s.evTotals, s.stream));
        // The gather needs the batch's totals only to size its buffers.  With the buffers a previous batch left it is enqueued
        // right here, behind the scan (every write is bounds-checked): tls_collect keeps its result when the totals fit and
        // redoes it after growing the buffers when they do not.  No host round trip between the scan and the gather.
        s.specGather = false;
        {
            const bool fusedLists = ctx->packed && s.mode == TLS_MODE_GENOME && (ctx->cfg.outputs & TLS_OUT_MATCHES);
            const uint64_t mCap = s.dMatches.cap / sizeof(DevMatch);
            uint64_t canCap = s.dCanIdx.cap / 4, nctCap = s.dNctIdx.cap / 4;
            if (fusedLists) { canCap = std::min<uint64_t>(canCap, s.dSelCan.cap / 4); nctCap = std::min<uint64_t>(nctCap, s.dSelNon.cap / 4); }
            if (mCap && canCap && nctCap && !ctx->noSpecGather) {
                if (int rc2 = enqueueGather(ctx, s, mCap, canCap, nctCap)) return rc2;
                s.specGather = true; s.specMatchCap = mCap; s.specSelCap = canCap; s.specNctCap = nctCap;
            }
        }
    } else {
        s.specGather = false;
        // a batch whose items are all empty: no kernel runs, every per-item counter reads as zero
        TLS_CUDA(cudaMemsetAsync(s.dTileBase.p, 0, 24, s.stream));
        std::memset(s.hScan.p, 0, 64);
        TLS_CUDA(s.hItemCounts.reserve(std::max<size_t>(s.nItems, 1) * sizeof(DevItemCounts)));
        std::memset(s.hItemCounts.p, 0, std::max<size_t>(s.nItems, 1) * sizeof(DevItemCounts));
    }
    return TLS_OK;
}

}  // namespace

extern "C" {

static int tls_submit_impl(tls_ctx *ctx, const tls_batch *b, int *ticket) {
    if (!ctx || !b || !ticket) return fail(TLS_ERR_INVALID, "null argument");
    if (b->n_items && (!b->seq || !b->offsets)) return fail(TLS_ERR_INVALID, "null batch arrays");
    if (b->mode > TLS_MODE_READS) return fail(TLS_ERR_INVALID, "bad mode %u", b->mode);
    if (b->mode == TLS_MODE_GENOME && !ctx->overlapRefusal.empty()) return fail(TLS_ERR_INVALID, "%s", ctx->overlapRefusal.c_str());
    int si = -1;
    for (int i = 0; i < 2; ++i) if (!ctx->slot[i].busy) { si = i; break; }
    if (si < 0) return fail(TLS_ERR_STATE, "two tickets already in flight");
    TLS_CUDA(cudaSetDevice(ctx->device));
    Slot &s = ctx->slot[si];
    Trace tr("submit");
    const uint32_t n = b->n_items;
    const uint32_t T = ctx->cfg.terminal_limit, S = ctx->cfg.step;
    s.nItems = n; s.mode = b->mode; s.collected = false;
    int rc = planBatch(ctx, b, s);
    if (rc) return rc;
    if (ctx->packed && b->mode == TLS_MODE_GENOME)
        for (uint32_t i = 0; i < n; ++i)
            if (s.itemLen[i] >= 0x80000000u) return fail(TLS_ERR_INVALID, "TLS_OUT_PACKED: item %u has 2^31 bases or more (tls_match4 holds 31-bit offsets)", i);
    const uint64_t base = n ? b->offsets[0] : 0;
    s.seqBytes = n ? b->offsets[n] - base : 0;

    // ---- stage the sequence -----------------------------------------------------------------
    if (b->seq_on_device) {
        s.devSeq = b->seq;
    } else {
        TLS_CUDA(s.dSeq.reserve(s.seqBytes + 512));
        cudaPointerAttributes attr{};
        const bool pinned = cudaPointerGetAttributes(&attr, b->seq) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const uint8_t *src = b->seq + base;
        if (!pinned && s.seqBytes) {
            // kRing worker threads on a large batch (one ring buffer each; chunk c belongs to thread c mod kRing), the
            // calling thread alone on a small one: a single core's memcpy (~10 GB/s) is slower than the DMA it feeds
            for (int k = 0; k < tls_ctx::kRing; ++k) {
                if (ctx->ringEv[k]) continue;
                TLS_CUDA(ctx->ring[k].reserve(tls_ctx::kRingBytes));
                TLS_CUDA(cudaEventCreateWithFlags(&ctx->ringEv[k], cudaEventDisableTiming));
                TLS_CUDA(cudaEventRecord(ctx->ringEv[k], s.stream));
            }
            const uint64_t nChunks = (s.seqBytes + tls_ctx::kRingBytes - 1) / tls_ctx::kRingBytes;
            const int workers = nChunks >= 2 * tls_ctx::kRing ? tls_ctx::kRing : 1;
            cudaError_t werr[tls_ctx::kRing] = {};
            auto work = [&](int w) {
                cudaError_t e = cudaSetDevice(ctx->device);
                for (uint64_t c = static_cast<uint64_t>(w); e == cudaSuccess && c < nChunks; c += static_cast<uint64_t>(workers)) {
                    const int k = workers == 1 ? static_cast<int>(c % tls_ctx::kRing) : w;
                    const uint64_t off = c * tls_ctx::kRingBytes;
                    const size_t nb = static_cast<size_t>(std::min<uint64_t>(tls_ctx::kRingBytes, s.seqBytes - off));
                    if ((e = cudaEventSynchronize(ctx->ringEv[k])) != cudaSuccess) break;      // the DMA that last read this buffer
                    std::memcpy(ctx->ring[k].p, src + off, nb);
                    if ((e = cudaMemcpyAsync(s.dSeq.as<uint8_t>() + off, ctx->ring[k].p, nb, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) break;
                    e = cudaEventRecord(ctx->ringEv[k], s.stream);
                }
                werr[w] = e;
            };
            if (workers == 1) work(0);
            else {
                std::vector<std::thread> pool;
                for (int w = 0; w < workers; ++w) pool.emplace_back(work, w);
                for (auto &t : pool) t.join();
            }
            for (int w = 0; w < workers; ++w) TLS_CUDA(werr[w]);
        } else if (s.seqBytes) {
            TLS_CUDA(cudaMemcpyAsync(s.dSeq.p, src, s.seqBytes, cudaMemcpyHostToDevice, s.stream));
        }
        TLS_CUDA(cudaMemsetAsync(s.dSeq.as<uint8_t>() + s.seqBytes, 0, 256, s.stream));
        s.devSeq = s.dSeq.as<uint8_t>();
    }
    tr.mark("stage_seq");
    const uint64_t devBase = b->seq_on_device ? 0 : base;   // offsets are relative to devSeq after this shift

    // ---- read mode: drop one trailing '\r' per read (src/read-filter.cpp:38-40) -------------------
    if (b->mode == TLS_MODE_READS && !b->seq_on_device) {
        for (uint32_t i = 0; i < n; ++i)
            if (s.itemLen[i] && b->seq[b->offsets[i + 1] - 1] == '\r') --s.itemLen[i];
    }

    // ---- reads: one warp per read, nothing but the pass flags comes back (scan_reads.cu) -----------------
    s.fusedReads = b->mode == TLS_MODE_READS && readConfigOk(ctx->dcfg) && !ctx->forceGeneral && !ctx->noReadKernel;
    if (s.fusedReads) {
        s.itemInfo.assign(n, DevItemInfo{});
        s.nBases = 0;
        for (uint32_t i = 0; i < n; ++i) { s.itemInfo[i].bufStart = b->offsets[i] - devBase; s.nBases += s.itemLen[i]; }
        s.nWindows = 0; s.winOff.assign(n + 1, 0); s.nRowItems = 0; s.nRowUnits = 0; s.nTilesAll = 0;
        const size_t bInfo = static_cast<size_t>(n) * sizeof(DevItemInfo), bLen = static_cast<size_t>(n) * 4;
        TLS_CUDA(s.hMeta3.reserve(bInfo + bLen + 64));
        uint8_t *h3 = s.hMeta3.as<uint8_t>();
        if (n) { std::memcpy(h3, s.itemInfo.data(), bInfo); std::memcpy(h3 + bInfo, s.itemLen.data(), bLen); }
        TLS_CUDA(s.dItemInfo.reserve(std::max<size_t>(bInfo, 16)));
        TLS_CUDA(s.dItemLen.reserve(std::max<size_t>(bLen, 16)));
        TLS_CUDA(s.dPass.reserve(std::max<size_t>(n, 1)));
        TLS_CUDA(s.hPass.reserve(std::max<size_t>(n, 1)));
        TLS_CUDA(cudaStreamWaitEvent(s.stream, s.evCopyDone, 0));
        if (n) {
            TLS_CUDA(cudaMemcpyAsync(s.dItemInfo.p, h3, bInfo, cudaMemcpyHostToDevice, s.stream));
            TLS_CUDA(cudaMemcpyAsync(s.dItemLen.p, h3 + bInfo, bLen, cudaMemcpyHostToDevice, s.stream));
        }
        TLS_CUDA(cudaMemsetAsync(s.dCounters.p, 0, 128, s.stream));      // [32] work counter, [36] status, [40] match total (u64)
        TLS_CUDA(cudaEventRecord(s.evAll0, s.stream));
        s.launches = 0;
        if (n && b->seq_on_device) {
            TLS_CUDA(launchTrimCrItems(s.devSeq, s.dItemInfo.as<DevItemInfo>(), n, s.dItemLen.as<uint32_t>(), s.stream));
            s.launches++;
        }
        TLS_CUDA(cudaEventRecord(s.evScan0, s.stream));
        if (n) {
            ReadParams rp{};
            rp.seq = s.devSeq; rp.items = s.dItemInfo.as<DevItemInfo>(); rp.itemLen = s.dItemLen.as<uint32_t>(); rp.nItems = n;
            rp.index8 = ctx->dIndex.as<uint8_t>();
            rp.flagTab = ctx->dFlags.as<uint8_t>() + ctx->dcfg.cls[0].flagOffset;
            rp.flagTabBytes = ((1u << (2 * ctx->dcfg.cls[0].len)) + 15u) & ~15u;
            rp.patLen = ctx->dcfg.cls[0].len;
            rp.pass = s.dPass.as<uint8_t>();
            rp.cursor = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 32);
            rp.status = reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 36);
            rp.matchTotal = reinterpret_cast<unsigned long long *>(s.dCounters.as<uint8_t>() + 40);
            rp.cfg = ctx->dcfg; rp.cfg.mode = TLS_MODE_READS;
            const int grid = static_cast<int>(std::min<uint32_t>(static_cast<uint32_t>(ctx->smCount), (n + (kReadThreads / 32) - 1) / (kReadThreads / 32)));
            TLS_CUDA(launchReadScan(rp, grid, s.stream));
            s.launches++;
        }
        TLS_CUDA(cudaEventRecord(s.evScan1, s.stream));
        TLS_CUDA(cudaEventRecord(s.evFast1, s.stream));
        TLS_CUDA(cudaEventRecord(s.evAll1, s.stream));
        if (n) TLS_CUDA(cudaMemcpyAsync(s.hPass.p, s.dPass.p, n, cudaMemcpyDeviceToHost, s.stream));
        TLS_CUDA(cudaMemcpyAsync(s.hCounters.p, s.dCounters.p, 64, cudaMemcpyDeviceToHost, s.stream));
        tr.mark("reads_enqueue");
        s.busy = true;
        *ticket = si;
        return TLS_OK;
    }

    // ---- regions ------------------------------------------------------------------------------
    uint64_t nWin = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t a = b->offsets[i] - devBase;
        const uint32_t len = s.itemLen[i];
        s.winOff[i] = nWin;
        auto add = [&](uint32_t rs, uint32_t rl) {
            DevRegion r{};
            r.bufStart = a + rs; r.len = rl; r.item = i; r.relStart = rs; r.itemLen = len;
            r.winBase = static_cast<uint32_t>(s.winOff[i]);
            s.regions.push_back(r);
            s.nBases += rl;
        };
        if (len == 0) continue;
        if (b->mode == TLS_MODE_GENOME) {
            nWin += (static_cast<uint64_t>(len) + S - 1) / S;          // src/teloscope.cpp:608-637
            add(0, len);
        } else if (b->mode == TLS_MODE_TIPS && len > static_cast<uint32_t>(2u * T)) {   // :576 (uint32 product)
            add(0, T);
            add(len - T, T);
        } else {
            add(0, len);
        }
    }
    s.winOff[n] = nWin;
    s.nWindows = nWin;
    if (nWin > 0xFFFFFFFFull) return fail(TLS_ERR_INVALID, "more than 2^32 windows in one batch");

    // ---- tiles ------------------------------------------------------------------------------------
    // Genome mode with a window-row configuration: every contig of >= kRowMinWindows windows becomes
    // tiles of 32 whole windows (one window per tile row).  Everything else: runs of (nearly) adjacent
    // regions are cut into pieces of <= kTileMaxLen bytes; of those the lean kernel takes the tiles that
    // lie inside one region, the general kernel the rest.
    tr.mark("plan_regions");
    s.tiles.clear(); s.tileIndex.clear(); s.nTilesAll = 0; s.rowRegions.clear(); s.nRowUnits = 0; s.nRowTiles = 0;
    {
        DevConfig probe = ctx->dcfg;
        probe.mode = b->mode;
        const bool rowOk = rowConfigOk(probe) && !ctx->forceGeneral && !ctx->noRowKernel;
        const uint32_t W = ctx->cfg.window_size;
        const uint64_t rowMinLen = static_cast<uint64_t>(W) * (ctx->rowMinWindows - 1) + 1;   // at least rowMinWindows windows
        size_t r = 0;
        const size_t nr = s.regions.size();
        while (r < nr) {
            const DevRegion &reg = s.regions[r];
            if (rowOk && reg.len >= rowMinLen) {
                const uint32_t nWinReg = (reg.len + W - 1) / W;
                const uint32_t nTilesReg = (nWinReg + kTileRows - 1) / kTileRows;
                DevRowRegion rr{};
                rr.bufStart = reg.bufStart; rr.len = reg.len; rr.relStart = reg.relStart; rr.itemLen = reg.itemLen;
                rr.winBase = reg.winBase; rr.item = reg.item;
                rr.tile0 = s.nTilesAll;
                rr.unit0 = s.nRowUnits;
                rr.nUnits = (nWinReg + kRowsPerUnit - 1) / kRowsPerUnit;      // kTileRows is a multiple of kRowsPerUnit
                s.rowRegions.push_back(rr);
                s.nRowUnits += rr.nUnits;
                // these tiles have no tile table entry: no kernel reads one (only lean / general tiles are looked up)
                s.nTilesAll += nTilesReg;
                s.nRowTiles += nTilesReg;
                ++r;
                continue;
            }
            uint64_t spanStart = reg.bufStart, spanEnd = spanStart + reg.len;
            size_t e = r + 1;
            while (e < nr && !(rowOk && s.regions[e].len >= rowMinLen) && s.regions[e].bufStart - spanEnd <= 4096) {
                spanEnd = s.regions[e].bufStart + s.regions[e].len;
                ++e;
            }
            size_t fr = r;
            for (uint64_t p = spanStart; p < spanEnd; p += kTileMaxLen) {
                DevTile t{};
                t.bufStart = p;
                t.len = static_cast<uint32_t>(std::min<uint64_t>(kTileMaxLen, spanEnd - p));
                while (fr + 1 < e && s.regions[fr].bufStart + s.regions[fr].len <= p) ++fr;
                t.firstRegion = static_cast<uint32_t>(fr);
                s.tiles.push_back(t);
                s.tileIndex.push_back(s.nTilesAll++);
            }
            r = e;
        }
    }

    // ---- per-item info for the kernels behind the scan; the windows the window-row kernel does not write ----
    s.itemInfo.assign(n, DevItemInfo{});
    for (uint32_t i = 0; i < n; ++i) {
        s.itemInfo[i].bufStart = b->offsets[i] - devBase;
        s.itemInfo[i].winBase = static_cast<uint32_t>(s.winOff[i]);
    }
    for (const DevRowRegion &rr : s.rowRegions) s.itemInfo[rr.item].rowItem = 1u;
    s.nRowItems = static_cast<uint32_t>(s.rowRegions.size());
    s.winRanges.clear(); s.winRangePrefix.assign(1, 0u);
    if (b->mode == TLS_MODE_GENOME && s.nRowItems) {
        for (uint32_t i = 0; i < n; ++i) {
            if (s.itemInfo[i].rowItem || s.winOff[i + 1] == s.winOff[i]) continue;
            const uint32_t x = static_cast<uint32_t>(s.winOff[i]), y = static_cast<uint32_t>(s.winOff[i + 1]);
            if (!s.winRanges.empty() && s.winRanges.back().y == x) s.winRanges.back().y = y;
            else s.winRanges.push_back(make_uint2(x, y));
        }
        for (const uint2 &r : s.winRanges) s.winRangePrefix.push_back(s.winRangePrefix.back() + (r.y - r.x));
    }
    tr.mark("plan_tiles");
    // ---- which kernel takes the other tiles: the lean kernel wants tiles inside one region -------------
    s.fastTiles.clear(); s.generalTiles.clear();
    s.fastTiles.reserve(s.tiles.size());
    {
        DevConfig probe = ctx->dcfg;
        probe.mode = b->mode;
        const bool fastOk = fastConfigOk(probe) && !ctx->forceGeneral;
        const uint32_t W = ctx->cfg.window_size;
        for (size_t tk = 0; tk < s.tiles.size(); ++tk) {
            const uint32_t ti = s.tileIndex[tk];
            const DevTile &t = s.tiles[tk];
            const DevRegion &r = s.regions[t.firstRegion];
            const bool single = r.bufStart <= t.bufStart && t.bufStart + t.len <= r.bufStart + r.len;
            if (!(fastOk && single)) { s.generalTiles.push_back(ti); continue; }
            DevFastTile f{};
            f.bufStart = t.bufStart; f.len = t.len; f.tile = ti;
            f.relBase = r.relStart + static_cast<uint32_t>(t.bufStart - r.bufStart);
            f.regionEndQ = r.relStart + r.len - f.relBase;
            f.item = r.item; f.itemLen = r.itemLen;
            f.firstWin = r.winBase + f.relBase / W;
            f.winDelta = r.winBase - f.firstWin;                   // mod 2^32
            const uint32_t termEnd = r.itemLen > T ? r.itemLen - T : 0;
            // no position of the tile is <= T or >= n - T  ->  none of its matches is "terminal"
            f.termMixed = !(f.relBase > T && static_cast<uint64_t>(f.relBase) + f.len <= termEnd);
            s.fastTiles.push_back(f);
        }
    }
    tr.mark("plan");
    // ---- upload metadata: ONE pinned block, ONE copy; the device buffers of the tables are views into it --------------
    {
        struct Part { const void *src; size_t bytes; DevBuf *view; size_t off; };
        Part parts[] = {
            {s.rowRegions.data(), s.rowRegions.size() * sizeof(DevRowRegion), &s.dRowRegions, 0},
            {s.fastTiles.data(), s.fastTiles.size() * sizeof(DevFastTile), &s.dFastTiles, 0},
            {s.generalTiles.data(), s.generalTiles.size() * 4, &s.dGeneralTiles, 0},
            {s.itemInfo.data(), static_cast<size_t>(n) * sizeof(DevItemInfo), &s.dItemInfo, 0},
            {s.winRanges.data(), s.winRanges.size() * sizeof(uint2), &s.dWinRanges, 0},          // followed directly by the prefix
            {s.winRangePrefix.data(), s.winRanges.empty() ? 0 : s.winRangePrefix.size() * 4, nullptr, 0},
            {s.absPos.data(), static_cast<size_t>(n) * 8, &s.dAbsPos, 0},
            {s.regions.data(), s.regions.size() * sizeof(DevRegion), &s.dRegions, 0},
            {s.itemLen.data(), static_cast<size_t>(n) * 4, &s.dItemLen, 0},
            {s.tiles.data(), s.tiles.size() * sizeof(DevTile), nullptr, 0},                      // compact: scattered below
            {s.tileIndex.data(), s.tiles.size() * 4, nullptr, 0},
        };
        size_t total = 0;
        for (Part &pt : parts) {
            if (pt.view) total = (total + 255) & ~size_t(255);       // a table starts on a 256-byte boundary; what follows a table directly is 8-aligned
            else total = (total + 7) & ~size_t(7);
            pt.off = total;
            total += pt.bytes;
        }
        total += 256;
        TLS_CUDA(s.hMeta.reserve(total));
        TLS_CUDA(s.dMetaAll.reserve(total));
        uint8_t *hm = s.hMeta.as<uint8_t>(), *dm = s.dMetaAll.as<uint8_t>();
        for (const Part &pt : parts) {
            if (pt.bytes) std::memcpy(hm + pt.off, pt.src, pt.bytes);
            if (pt.view) pt.view->setView(dm + pt.off);
        }
        tr.mark("meta_pack");
        TLS_CUDA(cudaMemcpyAsync(dm, hm, total, cudaMemcpyHostToDevice, s.stream));
        tr.mark("meta_copy");
        TLS_CUDA(s.dRowTiles.reserve(std::max<size_t>(static_cast<size_t>(s.nRowUnits) * sizeof(DevRowUnit), 16)));
        TLS_CUDA(s.dTiles.reserve(std::max<size_t>(static_cast<size_t>(s.nTilesAll) * sizeof(DevTile), 16)));
        s.planLaunches = (s.rowRegions.empty() || s.nRowUnits == 0 ? 0u : 1u) + (s.tiles.empty() ? 0u : 1u);
        if (!s.rowRegions.empty())
            TLS_CUDA(launchRowPlan(s.dRowRegions.as<DevRowRegion>(), static_cast<uint32_t>(s.rowRegions.size()), ctx->cfg.window_size,
                                   s.dRowTiles.as<DevRowUnit>(), s.nRowUnits, s.stream));
        if (!s.tiles.empty())                                         // the lean / general kernels' tiles, to their places in the tile table
            TLS_CUDA(launchScatterTiles(reinterpret_cast<const DevTile *>(dm + parts[9].off), reinterpret_cast<const uint32_t *>(dm + parts[10].off),
                                        static_cast<uint32_t>(s.tiles.size()), s.dTiles.as<DevTile>(), s.stream));
    }
    TLS_CUDA(s.dItemCounts.reserve(std::max<size_t>(n, 1) * sizeof(DevItemCounts)));
    TLS_CUDA(s.dItemBegin.reserve((static_cast<size_t>(n) + 1) * 8));
    TLS_CUDA(s.dWindows.reserve(std::max<uint64_t>(nWin, 1) * 32));
    TLS_CUDA(s.dZeroLanes.reserve(std::max<uint64_t>(nWin, 1) * 4));
    TLS_CUDA(s.dBoundaries.reserve(std::max<size_t>(n, 1) * 8));
    TLS_CUDA(s.dPass.reserve(std::max<size_t>(n, 1)));

    TLS_CUDA(cudaEventRecord(s.evAll0, s.stream));
    if (b->mode == TLS_MODE_READS && b->seq_on_device)
        TLS_CUDA(launchTrimCr(s.devSeq, s.dRegions.as<DevRegion>(), static_cast<uint32_t>(s.regions.size()),
                              s.dItemLen.as<uint32_t>(), s.stream));
    tr.mark("alloc_meta");
    if ((rc = enqueueScan(ctx, s))) return rc;
    tr.mark("enqueue");
    s.busy = true;
    *ticket = si;
    return TLS_OK;
}

static int tls_collect_impl(tls_ctx *ctx, int ticket, tls_result *res) {
    if (!ctx || !res || ticket < 0 || ticket > 1 || !ctx->slot[ticket].busy)
        return fail(TLS_ERR_STATE, "bad ticket");
    TLS_CUDA(cudaSetDevice(ctx->device));
    Slot &s = ctx->slot[ticket];
    const uint32_t n = s.nItems;
    int rc;
    Trace tr("collect");
    if (!s.collected && s.fusedReads) {
        TLS_CUDA(cudaStreamSynchronize(s.stream));
        const uint8_t *hc = s.hCounters.as<uint8_t>();
        if (*reinterpret_cast<const unsigned int *>(hc + 36) != 0)
            return fail(TLS_ERR_STATE, "read scan kernel could not lay out its shared memory; set TLS_NO_READ_KERNEL=1");
        TLS_CUDA(cudaEventElapsedTime(&s.scanMs, s.evScan0, s.evScan1));
        TLS_CUDA(cudaEventElapsedTime(&s.allMs, s.evAll0, s.evAll1));
        s.fastMs = 0;
        s.nMatches = *reinterpret_cast<const unsigned long long *>(hc + 40);
        s.terminalOff.assign(n + 1, 0); s.itsOff.assign(n + 1, 0); s.canOff.assign(n + 1, 0); s.nonOff.assign(n + 1, 0);
        s.terminal.clear(); s.its.clear();
        s.nCan.assign(n, 0); s.nFwd.assign(n, 0); s.nRev.assign(n, 0);
        s.pass.assign(n, 0);
        if (n) std::memcpy(s.pass.data(), s.hPass.p, n);
        s.nSelCan = s.nSelNon = 0;
        s.collected = true;
    }
    if (!s.collected) {
        // ---- wait for the scan; grow the overflow arena and rerun if it ran out --------------------
        for (;;) {
            if (s.nTilesAll) TLS_CUDA(cudaEventSynchronize(s.evTotals));       // not the stream: a gather enqueued behind the scan goes on meanwhile
            else TLS_CUDA(cudaStreamSynchronize(s.stream));
            const volatile unsigned long long *hs = s.hScan.as<unsigned long long>();   // ovCursor, totals[3], redo | status << 32
            if ((hs[4] >> 32) != 0)
                return fail(TLS_ERR_STATE, "lean scan kernel could not lay out its shared memory (window base %% 64 KiB); "
                                           "set TLS_FORCE_GENERAL=1");
            if (tr.on) fprintf(stderr, " row_tiles=%zu lean_tiles=%zu/%zu redo_rows=%u", s.nRowTiles, s.fastTiles.size(), static_cast<size_t>(s.nTilesAll),
                               static_cast<uint32_t>(hs[4]));
            const uint64_t ovUsed = hs[0];                     // 64-bit cursor: the true demand, never wrapped
            s.nMatches = hs[1]; s.nCanAll = hs[2]; s.nNctAll = hs[3];
            if (ovUsed <= s.ovCap) break;
            if (ovUsed + ovUsed / 4 > 0xF0000000ull)
                return fail(TLS_ERR_INVALID, "the batch needs %llu overflow match records (> 2^32); submit smaller batches",
                            static_cast<unsigned long long>(ovUsed));
            s.ovCap = static_cast<uint32_t>(ovUsed + ovUsed / 4);
            if ((rc = enqueueScan(ctx, s))) return rc;
        }
        if (s.nMatches > 0xFFFFFFF0ull) return fail(TLS_ERR_INVALID, "more than 2^32 matches in one batch; submit smaller batches");
        tr.mark("wait_scan");
        if (s.nTilesAll) {
            TLS_CUDA(cudaEventElapsedTime(&s.scanMs, s.evScan0, s.evScan1));
            TLS_CUDA(cudaEventElapsedTime(&s.fastMs, s.evScan0, s.evFast1));
        } else s.scanMs = s.fastMs = 0;
        const bool genome = s.mode == TLS_MODE_GENOME, reads = s.mode == TLS_MODE_READS;
        const bool wantSel = genome && (ctx->cfg.outputs & TLS_OUT_MATCHES) && s.nMatches > 0;

        // ---- slots -> one ordered list, blocks, -m lists: one enqueue, one synchronisation ------------------
        s.matchCap = std::max<uint64_t>(s.matchCap, s.nMatches);
        TLS_CUDA(s.dMatches.reserve(std::max<uint64_t>(s.matchCap, 1) * sizeof(DevMatch)));
        TLS_CUDA(s.dCanIdx.reserve(std::max<uint64_t>(s.nCanAll, 1) * 4));
        TLS_CUDA(s.dNctIdx.reserve(std::max<uint64_t>(s.nNctAll, 1) * 4));
        uint32_t nBlocks = 0;
        s.blockCap = std::max<uint32_t>(s.blockCap, 4096);
        const uint32_t runCap = static_cast<uint32_t>(std::min<uint64_t>(s.nCanAll / 4 + 64, 0x7FFFFFFFull));
        TLS_CUDA(s.dRunStarts.reserve(static_cast<size_t>(runCap) * 16));       // run starts (list walk) | chain starts (window-row items)
        if (wantSel) {
            TLS_CUDA(s.dSelCan.reserve(std::max<uint64_t>(s.nCanAll, 1) * sizeof(tls_match)));
            TLS_CUDA(s.dSelNon.reserve(std::max<uint64_t>(s.nNctAll, 1) * sizeof(tls_match)));
        }
        for (bool first = true;; first = false) {
            TLS_CUDA(s.dBlocks.reserve(static_cast<size_t>(s.blockCap) * sizeof(DevBlock)));
            uint8_t *cnt = s.dCounters.as<uint8_t>();
            TLS_CUDA(cudaMemsetAsync(cnt + 16, 0, 12, s.stream));           // [16] block count, [20] ITS run count, [24] ITS chain count
            if (first) {
                if (tr.on) for (cudaEvent_t &e : s.evDbg) if (!e) TLS_CUDA(cudaEventCreate(&e));
                const bool kept = s.specGather && s.nMatches <= s.specMatchCap && s.nCanAll <= s.specSelCap && s.nNctAll <= s.specNctCap;
                if (tr.on) fprintf(stderr, " gather=%s", kept ? "behind the scan" : (s.specGather ? "redone" : "here"));
                if (!kept && (rc = enqueueGather(ctx, s, s.matchCap, s.nCanAll, s.nNctAll))) return rc;
            }
            if (reads && n) TLS_CUDA(cudaMemsetAsync(s.dPass.p, 0, n, s.stream));
            if (n) {
                if (first && wantSel) {
                    // the -m lists only need the ordered list: final records and their copy home run on their own stream
                    // beside the block kernels
                    TLS_CUDA(cudaStreamWaitEvent(s.selStream, s.evGather, 0));
                    const size_t selRec = ctx->packed ? sizeof(tls_match4) : (ctx->compact ? sizeof(tls_match8) : sizeof(tls_match));
                    if (ctx->packed) {
                        // the gather kernels wrote the tls_match4 lists themselves: only the copies are left
                    } else if (ctx->compact) {
                        TLS_CUDA(launchSelectFinalize8(s.dMatches.as<DevMatch>(), s.dCanIdx.as<uint32_t>(), s.nCanAll, s.dSelCan.p, s.selStream));
                        TLS_CUDA(launchSelectFinalize8(s.dMatches.as<DevMatch>(), s.dNctIdx.as<uint32_t>(), s.nNctAll, s.dSelNon.p, s.selStream));
                    } else {
                        TLS_CUDA(launchSelectFinalize(s.dMatches.as<DevMatch>(), s.dCanIdx.as<uint32_t>(), s.nCanAll,
                                                      s.dItemBegin.as<unsigned long long>(), s.dAbsPos.as<unsigned long long>(), n,
                                                      s.dSelCan.p, s.selStream));
                        TLS_CUDA(launchSelectFinalize(s.dMatches.as<DevMatch>(), s.dNctIdx.as<uint32_t>(), s.nNctAll,
                                                      s.dItemBegin.as<unsigned long long>(), s.dAbsPos.as<unsigned long long>(), n,
                                                      s.dSelNon.p, s.selStream));
                    }
                    if (!ctx->packed) s.launches += (s.nCanAll ? 1 : 0) + (s.nNctAll ? 1 : 0);
                    if (s.nCanAll) {
                        TLS_CUDA(s.hSelCan.reserve(s.nCanAll * selRec));
                        TLS_CUDA(cudaMemcpyAsync(s.hSelCan.p, s.dSelCan.p, s.nCanAll * selRec, cudaMemcpyDeviceToHost, s.selStream));
                    }
                    if (s.nNctAll) {
                        TLS_CUDA(s.hSelNon.reserve(s.nNctAll * selRec));
                        TLS_CUDA(cudaMemcpyAsync(s.hSelNon.p, s.dSelNon.p, s.nNctAll * selRec, cudaMemcpyDeviceToHost, s.selStream));
                    }
                    if (tr.on && s.evDbg[2]) TLS_CUDA(cudaEventRecord(s.evDbg[2], s.selStream));
                }
                if (first && genome && ctx->dcfg.overlapWindows && s.nWindows) {
                    // step < window: the window table comes from the ordered list (block_kernels.cu), then goes home
                    TLS_CUDA(launchOverlapWindows(s.devSeq, s.dRegions.as<DevRegion>(), static_cast<uint32_t>(s.regions.size()),
                                                  s.dMatches.as<DevMatch>(), s.dItemBegin.as<unsigned long long>(),
                                                  static_cast<uint32_t>(s.nWindows), ctx->dcfg, s.dWindows.as<uint32_t>(), s.stream));
                    s.launches++;
                    if (ctx->compact) {
                        const size_t rec = ctx->packed ? 8 : 16;
                        TLS_CUDA(s.dWindows16.reserve(s.nWindows * 16));
                        if (ctx->packed) TLS_CUDA(launchPackWindows8(s.dWindows.as<uint32_t>(), s.dWindows16.p, s.nWindows, ctx->dcfg.cls[0].len, s.dWinExc.p,
                                                                     reinterpret_cast<unsigned int *>(s.dCounters.as<uint8_t>() + 48), s.stream));
                        else TLS_CUDA(launchPackWindows(s.dWindows.as<uint32_t>(), s.dWindows16.p, s.nWindows, s.stream));
                        s.launches++;
                        TLS_CUDA(s.hWindows.reserve(s.nWindows * rec));
                        TLS_CUDA(cudaMemcpyAsync(s.hWindows.p, s.dWindows16.p, s.nWindows * rec, cudaMemcpyDeviceToHost, s.stream));
                    } else {
                        TLS_CUDA(s.hWindows.reserve(s.nWindows * 32));
                        TLS_CUDA(cudaMemcpyAsync(s.hWindows.p, s.dWindows.p, s.nWindows * 32, cudaMemcpyDeviceToHost, s.stream));
                    }
                }
                BlockParams bp{};
                bp.matches = s.dMatches.as<DevMatch>();
                bp.itemBegin = s.dItemBegin.as<unsigned long long>();
                bp.itemCounts = s.dItemCounts.as<DevItemCounts>();
                bp.itemLen = s.dItemLen.as<uint32_t>();
                bp.nItems = n;
                if (s.nRowItems) {                          // window-row items: sparse record lists, walks may go back to the sequence
                    bp.items = s.dItemInfo.as<DevItemInfo>();
                    bp.seq = s.devSeq;
                    bp.flagTab = ctx->dFlags.as<uint8_t>() + ctx->dcfg.cls[0].flagOffset;
                    bp.patLen = ctx->dcfg.cls[0].len;
                    bp.zeroLanes = s.dZeroLanes.as<uint32_t>();
                }
                bp.blocks = s.dBlocks.as<DevBlock>();
                bp.blockCap = reads ? 0 : s.blockCap;       // read mode only needs the pass flag
                bp.blockCount = reinterpret_cast<unsigned int *>(cnt + 16);
                bp.fwdBoundary = s.dBoundaries.as<uint32_t>();
                bp.revBoundary = s.dBoundaries.as<uint32_t>() + n;
                bp.pass = s.dPass.as<uint8_t>();
                bp.cfg = ctx->dcfg;
                bp.cfg.mode = s.mode;
                TLS_CUDA(launchTerminal(bp, s.stream));
                s.launches++;
                if (tr.on && s.evDbg[1]) TLS_CUDA(cudaEventRecord(s.evDbg[1], s.stream));
                if (genome && s.nMatches >= 2 && s.nCanAll >= 4) {
                    const bool lists = s.nRowItems < n, rows = s.nRowItems != 0;
                    if (rows) TLS_CUDA(s.dConn.reserve(s.nCanAll + 64));     // itsWalkKernel reads 32 links at a time (before any launch: may synchronise)
                    // the two kinds of items are independent (blocks are appended with atomics): side by side when both exist
                    cudaStream_t listStream = s.stream;
                    if (lists && rows) {
                        TLS_CUDA(cudaEventRecord(s.evTerminal, s.stream));
                        TLS_CUDA(cudaStreamWaitEvent(s.itsStream, s.evTerminal, 0));
                        listStream = s.itsStream;
                    }
                    if (lists) {                            // items with complete record lists: runs are read from the list
                        TLS_CUDA(launchIts(bp, s.dCanIdx.as<uint32_t>(), s.nCanAll, s.dRunStarts.as<unsigned long long>(),
                                           reinterpret_cast<unsigned int *>(cnt + 20), runCap, listStream));
                        s.launches += 2;
                    }
                    if (rows) {                             // window-row items: connectivity from the zero-lane masks, runs from the sequence
                        TLS_CUDA(launchItsRows(bp, s.dCanIdx.as<uint32_t>(), s.nCanAll, s.dConn.as<uint8_t>(),
                                               s.dRunStarts.as<unsigned long long>() + runCap,
                                               reinterpret_cast<unsigned int *>(cnt + 24), runCap, s.stream));
                        s.launches += 3;
                    }
                    if (lists && rows) {
                        TLS_CUDA(cudaEventRecord(s.evItsLists, s.itsStream));
                        TLS_CUDA(cudaStreamWaitEvent(s.stream, s.evItsLists, 0));
                    }
                }
            }
            TLS_CUDA(cudaEventRecord(s.evAll1, s.stream));
            TLS_CUDA(cudaMemcpyAsync(s.hCounters.p, s.dCounters.p, 64, cudaMemcpyDeviceToHost, s.stream));
            TLS_CUDA(cudaStreamSynchronize(s.stream));
            nBlocks = *reinterpret_cast<unsigned int *>(s.hCounters.as<uint8_t>() + 16);
            if (reads || nBlocks <= s.blockCap) break;
            s.blockCap = nBlocks + nBlocks / 8 + 64;
        }
        if (reads) nBlocks = 0;
        if (tr.on && !reads)
            std::fprintf(stderr, "[tls] blocks %u, ITS runs from lists %u, ITS chains of window-row items %u, canonical %llu\n", nBlocks,
                         *reinterpret_cast<unsigned int *>(s.hCounters.as<uint8_t>() + 20),
                         *reinterpret_cast<unsigned int *>(s.hCounters.as<uint8_t>() + 24), static_cast<unsigned long long>(s.nCanAll));
        tr.mark("gather_blocks_select");
        unsigned long long selTotals[2] = {wantSel ? s.nCanAll : 0, wantSel ? s.nNctAll : 0};

        // ---- copy results back ------------------------------------------------------------------------
        s.nWinExc = (ctx->packed && !reads) ? *reinterpret_cast<unsigned int *>(s.hCounters.as<uint8_t>() + 48) : 0u;
        if (s.nWinExc) {
            TLS_CUDA(s.hWinExc.reserve(static_cast<size_t>(s.nWinExc) * sizeof(tls_window_exc)));
            TLS_CUDA(cudaMemcpyAsync(s.hWinExc.p, s.dWinExc.p, static_cast<size_t>(s.nWinExc) * sizeof(tls_window_exc), cudaMemcpyDeviceToHost, s.stream));
        }
        if (nBlocks) {
            TLS_CUDA(s.hBlocks.reserve(static_cast<size_t>(nBlocks) * sizeof(DevBlock)));
            TLS_CUDA(cudaMemcpyAsync(s.hBlocks.p, s.dBlocks.p, static_cast<size_t>(nBlocks) * sizeof(DevBlock),
                                     cudaMemcpyDeviceToHost, s.stream));
        }
        if (reads && n) {
            TLS_CUDA(s.hPass.reserve(n));
            TLS_CUDA(cudaMemcpyAsync(s.hPass.p, s.dPass.p, n, cudaMemcpyDeviceToHost, s.stream));
        }
        s.nSelCan = selTotals[0]; s.nSelNon = selTotals[1];
        TLS_CUDA(cudaStreamSynchronize(s.stream));
        TLS_CUDA(cudaStreamSynchronize(s.selStream));
        TLS_CUDA(cudaStreamSynchronize(s.copyStream));
        TLS_CUDA(cudaEventElapsedTime(&s.allMs, s.evAll0, s.evAll1));
        tr.mark("d2h");
        if (tr.on && s.evDbg[1] && !reads) {
            float tScan = 0, tGather = 0, tTerm = 0, tLists = 0;
            cudaEventElapsedTime(&tScan, s.evAll0, s.evScan1);
            cudaEventElapsedTime(&tGather, s.evAll0, s.evGather);
            cudaEventElapsedTime(&tTerm, s.evAll0, s.evDbg[1]);
            if (wantSel) cudaEventElapsedTime(&tLists, s.evAll0, s.evDbg[2]);
            cudaGetLastError();
            std::fprintf(stderr, "\n[tls] timeline from the first enqueue (ms): scan done %.3f, gather done %.3f, terminal done %.3f, blocks done %.3f, -m lists home %.3f\n",
                         tScan, tGather, tTerm, s.allMs, tLists);
        }

        // ---- assemble per-item views ----------------------------------------------------------------------
        s.terminalOff.assign(n + 1, 0); s.itsOff.assign(n + 1, 0); s.canOff.assign(n + 1, 0); s.nonOff.assign(n + 1, 0);
        s.terminal.clear(); s.its.clear();
        s.nCan.assign(n, 0); s.nFwd.assign(n, 0); s.nRev.assign(n, 0);
        s.pass.assign(n, 0);
        if (reads && n) std::memcpy(s.pass.data(), s.hPass.p, n);
        const DevItemCounts *ic = s.hItemCounts.as<DevItemCounts>();
        if (!reads)
            for (uint32_t i = 0; i < n; ++i) { s.nCan[i] = ic[i].can; s.nFwd[i] = ic[i].fwd; s.nRev[i] = ic[i].rev; }
        if (nBlocks) {
            std::vector<DevBlock> blk(s.hBlocks.as<DevBlock>(), s.hBlocks.as<DevBlock>() + nBlocks);
            // terminal blocks keep the reference's order (p walk then q walk, each in emission order);
            // ITS blocks are emitted by independent threads -> order by start
            std::sort(blk.begin(), blk.end(), [](const DevBlock &a, const DevBlock &b2) {
                if (a.item != b2.item) return a.item < b2.item;
                if (a.order != b2.order) return a.order < b2.order;
                return a.start < b2.start;
            });
            auto conv = [&](const DevBlock &d) {
                tls_block o{};
                o.start = s.absPos[d.item] + d.start; o.len = d.len; o.counts = d.counts; o.fwd_count = d.fwdCount;
                o.rev_count = d.counts - d.fwdCount; o.can_count = d.canCount; o.noncan_count = d.counts - d.canCount;
                o.total_cov = d.totalCov; o.fwd_cov = d.fwdCov; o.can_cov = d.canCov;
                o.has_valid_or = d.validOr; o.label = d.label;
                return o;
            };
            for (const DevBlock &d : blk) {
                if (d.kind == 1) { s.its.push_back(conv(d)); s.itsOff[d.item + 1]++; }
                else { s.terminal.push_back(conv(d)); s.terminalOff[d.item + 1]++; }
            }
        }
        for (uint32_t i = 0; i < n; ++i) { s.terminalOff[i + 1] += s.terminalOff[i]; s.itsOff[i + 1] += s.itsOff[i]; }
        if (wantSel) {
            // the device wrote final records; only the per-item offsets are left to the host
            uint64_t kc = 0, kn = 0;
            for (uint32_t i = 0; i < n; ++i) { s.canOff[i] = kc; s.nonOff[i] = kn; kc += ic[i].can; kn += ic[i].nonCanTerm; }
            s.canOff[n] = kc; s.nonOff[n] = kn;
            if (kc != s.nSelCan || kn != s.nSelNon)
                return fail(TLS_ERR_STATE, "internal: selected match lists disagree with the per-item counters");
        }
        tr.mark("assemble");
        s.collected = true;
    }
    std::memset(res, 0, sizeof *res);
    res->n_items = n; res->mode = s.mode;
    res->window_off = s.winOff.data();
    res->windows = ctx->compact ? nullptr : s.hWindows.as<tls_window>();
    res->windows16 = (ctx->compact && !ctx->packed) ? s.hWindows.as<tls_window16>() : nullptr;
    res->windows8 = ctx->packed ? s.hWindows.as<tls_window8>() : nullptr;
    res->n_window_exc = ctx->packed ? s.nWinExc : 0;
    res->window_exc = (ctx->packed && s.nWinExc) ? s.hWinExc.as<tls_window_exc>() : nullptr;
    res->canonical4 = s.nSelCan && ctx->packed ? s.hSelCan.as<tls_match4>() : nullptr;
    res->noncanonical4 = s.nSelNon && ctx->packed ? s.hSelNon.as<tls_match4>() : nullptr;
    res->pattern_len = ctx->dcfg.nClasses == 1 ? ctx->dcfg.cls[0].len : 0u;
    res->terminal_off = s.terminalOff.data(); res->terminal = s.terminal.data();
    res->its_off = s.itsOff.data(); res->its = s.its.data();
    res->canonical_off = s.canOff.data(); res->canonical = s.nSelCan && !ctx->compact ? s.hSelCan.as<tls_match>() : nullptr;
    res->noncanonical_off = s.nonOff.data(); res->noncanonical = s.nSelNon && !ctx->compact ? s.hSelNon.as<tls_match>() : nullptr;
    res->canonical8 = s.nSelCan && ctx->compact && !ctx->packed ? s.hSelCan.as<tls_match8>() : nullptr;
    res->noncanonical8 = s.nSelNon && ctx->compact && !ctx->packed ? s.hSelNon.as<tls_match8>() : nullptr;
    res->n_canonical = s.nCan.data(); res->n_fwd = s.nFwd.data(); res->n_rev = s.nRev.data();
    res->pass = s.pass.data();
    res->scan_kernel_ms = s.scanMs; res->device_ms = s.allMs;
    res->n_bases_scanned = s.nBases; res->n_matches = s.nMatches; res->n_kernel_launches = s.launches;
    res->lean_kernel_ms = s.fastMs;
    return TLS_OK;
}

int tls_release(tls_ctx *ctx, int ticket) {
    if (!ctx || ticket < 0 || ticket > 1 || !ctx->slot[ticket].busy) return fail(TLS_ERR_STATE, "bad ticket");
    ctx->slot[ticket].busy = false;
    return TLS_OK;
}

}  // extern "C"

// ---- the exported entry points: nothing throws across the ABI (std::bad_alloc / length_error from the host-side
//      containers become TLS_ERR_NOMEM / TLS_ERR_INVALID with a message) ----------------------------------------------
namespace {
template <typename F> int guarded(F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return fail(TLS_ERR_NOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return fail(TLS_ERR_INVALID, "internal error: %s", e.what());
    } catch (...) {
        return fail(TLS_ERR_INVALID, "internal error");
    }
}
}  // namespace

extern "C" {

int tls_expand_patterns(const char *raw_csv, int edit_distance, const char *canonical_fwd, char *text, size_t text_len,
                        uint8_t *is_forward, uint8_t *is_canonical, uint32_t cap, uint32_t *n_out, size_t *text_need) {
    return guarded([&] { return tls_expand_patterns_impl(raw_csv, edit_distance, canonical_fwd, text, text_len, is_forward, is_canonical, cap, n_out, text_need); });
}
int tls_create(const tls_config *cfg, int device, tls_ctx **out) {
    return guarded([&] { return tls_create_impl(cfg, device, out); });
}
int tls_submit(tls_ctx *ctx, const tls_batch *b, int *ticket) {
    return guarded([&] { return tls_submit_impl(ctx, b, ticket); });
}
int tls_collect(tls_ctx *ctx, int ticket, tls_result *res) {
    return guarded([&] { return tls_collect_impl(ctx, ticket, res); });
}

}  // extern "C"
