/*
 * teloscope_b200.h — C ABI of the B200-native telomeric motif scan.
 *
 * This is the drop-in boundary for the ONE hot path of vgl-hub/teloscope that runs on the GPU:
 *
 *     SegmentData Teloscope::scanSegment(std::string &sequence, uint64_t absPos, bool tipsOnly)
 *         reference: include/teloscope.h:260, src/teloscope.cpp:537-658
 *     bool ReadTelomereFilter::matches(std::string sequence)
 *         reference: include/read-filter.h:16, src/read-filter.cpp:37-45
 *     expandPatternsWithOrientation(rawPatterns, editDistance, canonicalFwd)
 *         reference: include/tools.h:45-48, src/tools.cpp:201-283   (host, feeds the device LUTs)
 *
 * The reference calls scanSegment once per contig / GFA segment end / read from thread-pool jobs
 * (src/input.cpp:842,892,977; src/read-filter.cpp:43).  Here the caller hands over a BATCH of
 * such sequences and gets back, per sequence, exactly what SegmentData holds: window counters,
 * terminal blocks, interstitial blocks, canonical / terminal non-canonical matches — or, in read
 * mode, the pass flag.  Plain pointers and sizes only; no C++ or torch types cross this line.
 *
 * Conventions: every function returns TLS_OK (0) or a negative tls_status and leaves a message
 * retrievable with tls_last_error(); nothing throws or exits across the ABI.  One ctx per device;
 * a ctx is used by one host thread at a time.  There is NO CPU fallback: tls_create fails when no
 * sm_100 device is present.
 */
#ifndef TELOSCOPE_B200_H
#define TELOSCOPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLS_ABI_VERSION 4   /* 2: tls_result grew (windows16, canonical8, noncanonical8), TLS_OUT_COMPACT; 3: TLS_OUT_PACKED (canonical4, noncanonical4, pattern_len);
                               4: packed windows are tls_window8 (windows8, window_exc, n_window_exc) */

typedef enum {
    TLS_OK = 0,
    TLS_ERR_INVALID = -1,    /* bad argument / config                                   */
    TLS_ERR_NO_DEVICE = -2,  /* no CUDA device of compute capability 10.x               */
    TLS_ERR_CUDA = -3,       /* CUDA runtime error (message has the details)            */
    TLS_ERR_NOMEM = -4,
    TLS_ERR_AMBIGUOUS = -5,  /* one pattern string produced with both orientation flags */
    TLS_ERR_STATE = -6       /* bad ticket / collect without submit                     */
} tls_status;

/* ---- records (layouts shared with the test oracle) ------------------------------------------ */

typedef struct {          /* MatchInfo, include/teloscope.h:89-95 */
    uint64_t pos;         /* absPos + offset in the sequence */
    uint16_t len;
    uint8_t is_forward;
    uint8_t is_canonical;
    uint32_t pad_;
} tls_match;              /* 16 bytes */

typedef struct {          /* TelomereBlock, include/teloscope.h:103-117 */
    uint64_t start;
    uint32_t len, counts, fwd_count, rev_count, can_count, noncan_count;
    uint32_t total_cov, fwd_cov, can_cov;
    uint8_t has_valid_or;
    char label;           /* 'p' / 'q' (terminal), 'p' / 'q' / 'b' (interstitial) */
    uint16_t pad_;
} tls_block;              /* 48 bytes */

typedef struct {          /* integer part of WindowData, include/teloscope.h:119-137 */
    uint32_t acgt[4];     /* A, C, G, T (0 unless TLS_OUT_ACGT)                */
    uint32_t can_cov, noncan_cov, fwd_cov, rev_cov;
} tls_window;             /* 32 bytes; start = abs_pos + i*step, size = min(window, n - i*step) */

/* Compact forms of the two bulk results (TLS_OUT_COMPACT): half the bytes per record over PCIe and in host memory.
 * The counters of a window are bounded by window_size x (sum of the pattern lengths), so they fit 16 bits whenever
 * that product does (tls_create refuses the flag otherwise); a match position relative to its item fits 32 bits. */
typedef struct {
    uint16_t acgt[4];
    uint16_t can_cov, noncan_cov, fwd_cov, rev_cov;
} tls_window16;           /* 16 bytes, same fields as tls_window */

typedef struct {
    uint32_t pos;         /* offset in the item: MatchInfo.position = abs_pos[item] + pos */
    uint16_t len;
    uint8_t is_forward;
    uint8_t is_canonical;
} tls_match8;             /* 8 bytes */

/* Packed forms (TLS_OUT_PACKED): the fewest bytes over the host link — what bounds a multi-GPU box whose GPUs share it.
 * Only for pattern tables of ONE length L and window_size <= 1023 (tls_create refuses the flag otherwise): the window
 * holds COUNTS of matches (covered bp = count x L). */
typedef struct {
    uint32_t w[2];        /* w[0] = A | C << 10 | G << 20 | (T & 3) << 30;
                             w[1] = T >> 2 | canonical matches << 8 | all matches << 16 | forward matches << 24.
                             A match count of 255 means "255 or more": the exact counts of such a window are in
                             tls_result.window_exc (no window of a telomere holds that many matches of one length >= 4).
                             can_cov = canonical x L, noncan_cov = (all - canonical) x L, fwd_cov = forward x L,
                             rev_cov = (all - forward) x L (L = tls_result.pattern_len) */
} tls_window8;            /* 8 bytes */
typedef struct {
    uint32_t window;      /* index into windows8 */
    uint32_t canonical, all, forward;
} tls_window_exc;         /* windows with a match count >= 255, in no particular order */
typedef uint32_t tls_match4;   /* bits 0..30: offset in the item (items of 2^31 bases or more are refused at tls_submit),
                                  bit 31: forward strand; len = pattern_len; canonical list: is_canonical = 1 */

/* ---- configuration ---------------------------------------------------------------------------- */

enum {                    /* tls_config.outputs */
    TLS_OUT_ACGT = 1u << 0,     /* -g / -e : per-window A,C,G,T counts (src/teloscope.cpp:418)  */
    TLS_OUT_MATCHES = 1u << 1,  /* -m      : canonical + terminal non-canonical match lists     */
    TLS_OUT_ITS = 1u << 2,      /* -i      : interstitial blocks are always computed in genome
                                             mode (the report needs their count); flag kept for
                                             symmetry with the reference's toggles              */
    TLS_OUT_COMPACT = 1u << 3,  /* windows come back as tls_window16 (tls_result.windows16) and the -m lists as
                                   tls_match8 (canonical8 / noncanonical8); windows / canonical / noncanonical
                                   are then NULL                                                */
    TLS_OUT_PACKED = 1u << 4    /* windows as tls_window8 (windows8 + window_exc), -m lists as tls_match4 (canonical4 /
                                   noncanonical4); every other window / match pointer is NULL      */
};

typedef struct {
    uint32_t abi_version;            /* TLS_ABI_VERSION */
    /* pattern table = userInput.patternInfo + isCanonical (include/teloscope.h:241-247) */
    uint32_t n_patterns;
    const char *const *patterns;     /* ACGT strings, the order expandPatternsWithOrientation gives */
    const uint8_t *is_forward;
    const uint8_t *is_canonical;
    /* UserInputTeloscope scan parameters (include/input.h:36-48) */
    /* 0 < step <= window_size.  step < window_size (overlapping windows, src/teloscope.cpp:409-434,
       485,512-529) is supported for pattern tables of ONE length L with step >= L and window_size -
       step >= L; otherwise the reference's uint32 scan start wraps (:413-415) or its match vectors
       are not position-ordered, and tls_submit refuses TLS_MODE_GENOME batches (other modes have no
       windows and run). */
    uint32_t window_size, step, terminal_limit;
    uint16_t max_match_dist, min_block_len, max_block_dist, min_block_counts;
    float min_block_density;
    uint32_t outputs;                /* TLS_OUT_* */
} tls_config;

typedef enum {
    TLS_MODE_GENOME = 0,  /* scanSegment(seq, absPos, tipsOnly=false): windows + ITS            */
    TLS_MODE_TIPS = 1,    /* scanSegment(seq, absPos, tipsOnly=true): ultra-fast / GFA mode     */
    TLS_MODE_READS = 2    /* ReadTelomereFilter::matches: tips scan with terminalLimit =
                             UINT32_MAX/2, trailing '\r' dropped, result = pass flag            */
} tls_mode;

typedef struct {
    const uint8_t *seq;        /* n_items sequences, item i = seq[offsets[i] .. offsets[i+1]);
                                  raw ASCII of any case — the scan folds case like the reference's
                                  callers do with unmaskSequence before scanSegment                */
    const uint64_t *offsets;   /* n_items + 1 entries                                              */
    const uint64_t *abs_pos;   /* per item, NULL = all zero                                        */
    uint32_t n_items;
    uint32_t mode;             /* tls_mode */
    uint32_t seq_on_device;    /* 1: seq is a DEVICE pointer on the ctx's device (16-byte aligned,
                                  readable up to offsets[n]+256 bytes); offsets/abs_pos stay host   */
    uint32_t pad_;
} tls_batch;

typedef struct {
    uint32_t n_items;
    uint32_t mode;
    /* all *_off arrays have n_items + 1 entries and index the record arrays below */
    const uint64_t *window_off;       const tls_window *windows;        /* genome mode           */
    const uint64_t *terminal_off;     const tls_block *terminal;        /* scan order, p then q  */
    const uint64_t *its_off;          const tls_block *its;             /* genome mode           */
    const uint64_t *canonical_off;    const tls_match *canonical;       /* TLS_OUT_MATCHES       */
    const uint64_t *noncanonical_off; const tls_match *noncanonical;    /* TLS_OUT_MATCHES       */
    const uint32_t *n_canonical;      /* per item: canonicalMatches.size() (genome mode)          */
    const uint32_t *n_fwd, *n_rev;    /* per item: fwdMatches.size(), revMatches.size()           */
    const uint8_t *pass;              /* read mode: 1 = terminalBlocks non-empty                  */
    /* measurement (CUDA events on the ctx's stream) */
    float scan_kernel_ms;             /* scan kernels (lean + general), CUDA events                */
    float device_ms;                  /* all kernels of the batch                                 */
    uint64_t n_bases_scanned;         /* bases staged through the scan kernel                     */
    uint64_t n_matches;               /* match records the scan emitted                           */
    uint32_t n_kernel_launches;
    float lean_kernel_ms;             /* lean scan kernel alone (0 when the batch has no lean tiles) */
    /* TLS_OUT_COMPACT (same *_off arrays) */
    const tls_window16 *windows16;
    const tls_match8 *canonical8, *noncanonical8;
    /* TLS_OUT_PACKED (same *_off arrays) */
    const tls_window8 *windows8;
    const tls_match4 *canonical4, *noncanonical4;
    uint32_t pattern_len;             /* the pattern length when the table has one, else 0 */
    uint32_t n_window_exc;            /* windows whose match counts did not fit tls_window8 */
    const tls_window_exc *window_exc;
} tls_result;

typedef struct tls_ctx tls_ctx;

/* ---- host side: pattern table ----------------------------------------------------------------- */

/* expandPatternsWithOrientation (src/tools.cpp:201-283) + isCanonical.  raw_csv: comma separated
 * seeds, already upper-cased (IUPAC allowed).  On success *n_out patterns are written as
 * NUL-terminated strings back to back into text (text_len bytes), flags into is_forward /
 * is_canonical (cap entries).  Call with text == NULL to query sizes (*n_out, *text_need). */
int tls_expand_patterns(const char *raw_csv, int edit_distance, const char *canonical_fwd,
                        char *text, size_t text_len, uint8_t *is_forward, uint8_t *is_canonical,
                        uint32_t cap, uint32_t *n_out, size_t *text_need);

/* ---- device side ------------------------------------------------------------------------------ */

int tls_device_count(void);                               /* sm_100 devices visible */
int tls_create(const tls_config *cfg, int device, tls_ctx **out);
void tls_destroy(tls_ctx *ctx);

/* Asynchronous: stages the batch (cudaMemcpyAsync), enqueues the kernels and the result copies on
 * the ctx's stream and returns a ticket.  At most two tickets may be in flight.
 * Lifetime of the caller's buffers: offsets / abs_pos are copied before tls_submit returns.  A
 * PAGEABLE seq buffer is staged through the library's pinned ring before tls_submit returns and may
 * be reused at once.  A PINNED host seq buffer (cudaHostAlloc / cudaHostRegister) is read by an
 * asynchronous DMA, and a DEVICE seq buffer (seq_on_device) is read by the kernels, including
 * those tls_collect enqueues: both must stay untouched until tls_collect(ticket) has returned. */
int tls_submit(tls_ctx *ctx, const tls_batch *batch, int *ticket);
/* Blocks until the ticket's work is done and fills *res with pointers into ctx-owned host memory,
 * valid until tls_release(ticket). */
int tls_collect(tls_ctx *ctx, int ticket, tls_result *res);
int tls_release(tls_ctx *ctx, int ticket);

const char *tls_last_error(void);                         /* thread-local message */

#ifdef __cplusplus
}
#endif
#endif /* TELOSCOPE_B200_H */
