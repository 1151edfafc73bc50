// teloscope-b200 --bam-subset: BAM records whose read has a Teloscope-valid telomeric block, written back verbatim as BAM.
//
// What the reference does (src/bam.cpp:188-316, src/bgzf.cpp:57-304), restated for a GPU scorer and host threads:
//   * BGZF in (src/bgzf.cpp:57-194): every block header is checked like BgzfReader::loadBlock does (gzip flags, the BC
//     subfield, optional FNAME / FCOMMENT / FHCRC, sizes), the payloads of a segment of blocks are inflated IN PARALLEL
//     (raw deflate, -15) and their CRC-32 / ISIZE verified.  Errors surface in stream order, when the BAM parser needs
//     the bytes of the failing block, with the reference's messages;
//   * BAM parse (src/bam.cpp:75-186): header copied through, records cut at block_size, SEQ located behind read name and
//     CIGAR and decoded from 4 bits ("=ACMGRSVTWYHKDBN") in parallel into TLS_MODE_READS batches for the C ABI
//     (ReadTelomereFilter::matches on the device: one warp per read, include/teloscope_b200.h);
//   * records without SEQ are counted and dropped, records whose read passes are appended to the output in input order;
//   * BGZF out (src/bgzf.cpp:225-304): the output byte stream is cut every 65 280 bytes, each piece deflated on its own with
//     deflateInit2(Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY), pieces that do not fit 64 KiB are halved,
//     fixed 18-byte header, CRC-32 + ISIZE footer, EOF marker block.  Pieces are independent, so they are deflated in parallel
//     and written in order: the file is byte-identical to the reference's with the same zlib.
// Memory is bounded: the input is walked in segments of blocks (about 256 MiB inflated), an incomplete record is carried over.
#include "pipeline.hpp"

#include <zlib.h>

#include <array>
#include <atomic>
#include <mutex>
#include <stdexcept>

namespace {

constexpr size_t kMaxBlockSize = 65536, kMaxUncompressed = 65280;
constexpr size_t kBamCoreSize = 32, kBamMaxHeaderText = 1ull << 30, kBamMaxReferenceName = 1ull << 20, kBamMaxRecordSize = 256ull << 20;
const uint8_t kEofBlock[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};

inline uint16_t getU16(const uint8_t *d) { return static_cast<uint16_t>(d[0] | (d[1] << 8)); }
inline uint32_t getU32(const uint8_t *d) { return static_cast<uint32_t>(d[0]) | (static_cast<uint32_t>(d[1]) << 8) | (static_cast<uint32_t>(d[2]) << 16) | (static_cast<uint32_t>(d[3]) << 24); }
inline int32_t getI32(const uint8_t *d) { const uint32_t v = getU32(d); int32_t r; std::memcpy(&r, &v, 4); return r; }
inline void putU16(uint8_t *d, uint16_t v) { d[0] = static_cast<uint8_t>(v); d[1] = static_cast<uint8_t>(v >> 8); }
inline void putU32(uint8_t *d, uint32_t v) { d[0] = static_cast<uint8_t>(v); d[1] = static_cast<uint8_t>(v >> 8); d[2] = static_cast<uint8_t>(v >> 16); d[3] = static_cast<uint8_t>(v >> 24); }

struct BamError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- BGZF reader over an in-memory (mapped) file ---------------------------------------------------------------------
struct BlockRef {
    size_t begin, total;          // the block in the input
    size_t payload, payloadLen;   // deflate payload
    uint32_t crc, isize;
    size_t dst;                   // offset of its bytes in the segment buffer
};

class BgzfIn {
public:
    BgzfIn(const uint8_t *p, size_t n, unsigned threads) : p_(p), n_(n), threads_(threads) {}
    bool physicalEof = false, sawEofBlock = false;

    // Appends the bytes of the next segment of blocks to `out`.  false: nothing more (end of input); throws the error of the
    // first bad block when no good block precedes it in this call.
    bool nextSegment(std::vector<uint8_t> &out, size_t wantBytes) {
        if (!pending_.empty()) { const std::string e = pending_; pending_.clear(); throw BamError(e); }
        if (physicalEof) return false;
        std::vector<BlockRef> blocks;
        size_t bytes = 0;
        std::string headerError;
        while (bytes < wantBytes) {
            BlockRef b{};
            const int rc = readHeader(b, headerError);
            if (rc <= 0) break;                               // 0: end of input, -1: header error (message kept)
            b.dst = bytes;
            bytes += b.isize;
            blocks.push_back(b);
        }
        const size_t base = out.size();
        out.resize(base + bytes);
        std::vector<std::string> errors(blocks.size());
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < blocks.size();) errors[i] = inflateBlock(blocks[i], out.data() + base + blocks[i].dst);
        };
        tlsb::runThreads(std::min<size_t>(threads_, std::max<size_t>(1, blocks.size() / 16)), work);
        size_t good = blocks.size();
        for (size_t i = 0; i < blocks.size(); ++i) if (!errors[i].empty()) { good = i; break; }
        for (size_t i = 0; i < good; ++i)
            if (blocks[i].isize == 0 && blocks[i].total == sizeof kEofBlock && !std::memcmp(p_ + blocks[i].begin, kEofBlock, sizeof kEofBlock)) sawEofBlock = true;
        if (good < blocks.size()) {                           // everything before the bad block is delivered first
            out.resize(base + blocks[good].dst);
            pending_ = errors[good];
            physicalEof = false;
            pos_ = n_;                                        // nothing is read behind an error
        } else if (!headerError.empty()) {
            pending_ = headerError;
        }
        if (out.size() == base) {
            if (!pending_.empty()) { const std::string e = pending_; pending_.clear(); throw BamError(e); }
            return !physicalEof;
        }
        return true;
    }

private:
    const uint8_t *p_;
    size_t n_, pos_ = 0;
    unsigned threads_;
    std::string pending_;

    // BgzfReader::loadBlock up to the payload (src/bgzf.cpp:57-149)
    int readHeader(BlockRef &b, std::string &err) {
        const size_t left = n_ - pos_;
        if (left == 0) { physicalEof = true; return 0; }
        if (left < 12) { err = "truncated BGZF header"; pos_ = n_; return -1; }
        const uint8_t *h = p_ + pos_;
        if (h[0] != 31 || h[1] != 139 || h[2] != 8) { err = "input is not BGZF-compressed BAM"; pos_ = n_; return -1; }
        const uint8_t flags = h[3];
        if ((flags & 0x04) == 0 || (flags & 0xe0) != 0) { err = "invalid BGZF gzip flags"; pos_ = n_; return -1; }
        const size_t extraLen = getU16(h + 10);
        if (left < 12 + extraLen) { err = "truncated BGZF extra field"; pos_ = n_; return -1; }
        const uint8_t *extra = h + 12;
        bool found = false;
        uint16_t stored = 0;
        for (size_t q = 0; q < extraLen;) {
            if (extraLen - q < 4) { err = "malformed BGZF extra field"; pos_ = n_; return -1; }
            const size_t sub = getU16(extra + q + 2), end = q + 4 + sub;
            if (end > extraLen) { err = "malformed BGZF extra subfield"; pos_ = n_; return -1; }
            if (extra[q] == 'B' && extra[q + 1] == 'C') {
                if (sub != 2 || found) { err = "invalid BGZF BC subfield"; pos_ = n_; return -1; }
                stored = getU16(extra + q + 4);
                found = true;
            }
            q = end;
        }
        if (!found) { err = "BGZF block is missing the BC subfield"; pos_ = n_; return -1; }
        const size_t total = static_cast<size_t>(stored) + 1, bytesRead = 12 + extraLen;
        if (total > kMaxBlockSize || total < bytesRead + 8) { err = "invalid BGZF block size"; pos_ = n_; return -1; }
        if (left < total) { err = "truncated BGZF block"; pos_ = n_; return -1; }
        size_t payload = bytesRead;
        const size_t footer = total - 8;
        auto skipZ = [&](const char *field) {
            while (payload < footer && h[payload] != 0) ++payload;
            if (payload == footer) { err = std::string("unterminated BGZF ") + field; return false; }
            ++payload;
            return true;
        };
        if ((flags & 0x08) && !skipZ("filename")) { pos_ = n_; return -1; }
        if ((flags & 0x10) && !skipZ("comment")) { pos_ = n_; return -1; }
        if (flags & 0x02) {
            if (footer - payload < 2) { err = "truncated BGZF header checksum"; pos_ = n_; return -1; }
            const uLong crc = crc32(crc32(0L, Z_NULL, 0), h, static_cast<uInt>(payload));
            if (static_cast<uint16_t>(crc) != getU16(h + payload)) { err = "BGZF header checksum mismatch"; pos_ = n_; return -1; }
            payload += 2;
        }
        b.begin = pos_; b.total = total; b.payload = pos_ + payload; b.payloadLen = footer - payload;
        b.crc = getU32(h + footer); b.isize = getU32(h + footer + 4);
        if (b.isize > kMaxBlockSize) { err = "BGZF uncompressed block is too large"; pos_ = n_; return -1; }
        pos_ += total;
        return 1;
    }

    // the rest of loadBlock (src/bgzf.cpp:151-181): "" or the reference's message
    std::string inflateBlock(const BlockRef &b, uint8_t *dst) const {
        uint8_t dummy[1];
        z_stream zs{};
        zs.next_in = const_cast<Bytef *>(p_ + b.payload);
        zs.avail_in = static_cast<uInt>(b.payloadLen);
        zs.next_out = b.isize ? dst : dummy;
        zs.avail_out = std::max<uInt>(b.isize, 1);
        if (inflateInit2(&zs, -15) != Z_OK) return "could not initialize BGZF decompressor";
        const int st = inflate(&zs, Z_FINISH);
        const size_t outBytes = zs.total_out, inBytes = zs.total_in;
        const int endSt = inflateEnd(&zs);
        if (st != Z_STREAM_END || outBytes != b.isize || inBytes != b.payloadLen) return "invalid BGZF deflate payload";
        if (endSt != Z_OK) return "could not finalize BGZF decompressor";
        const uLong crc = crc32(crc32(0L, Z_NULL, 0), b.isize ? dst : Z_NULL, b.isize);
        if (crc != b.crc) return "BGZF checksum mismatch";
        return std::string();
    }
};

// ---- BGZF writer: BgzfWriter (src/bgzf.cpp:196-304), pieces deflated in parallel ---------------------------------------
class BgzfOut {
public:
    BgzfOut(FILE *f, unsigned threads) : f_(f), threads_(threads) {}
    void write(const uint8_t *d, size_t n) {
        pending_.insert(pending_.end(), d, d + n);
        if (pending_.size() >= kFlushPieces * kMaxUncompressed) flush(false);
    }
    void finish() {
        flush(true);
        if (fwrite(kEofBlock, 1, sizeof kEofBlock, f_) != sizeof kEofBlock || fflush(f_) != 0) throw BamError("failed while finalizing BAM output");
    }

private:
    static constexpr size_t kFlushPieces = 512;
    FILE *f_;
    unsigned threads_;
    std::vector<uint8_t> pending_;

    static void compressRaw(const uint8_t *d, size_t n, std::vector<uint8_t> &out) {
        z_stream zs{};
        if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw BamError("could not initialize BGZF compressor");
        out.resize(deflateBound(&zs, n));
        zs.next_in = const_cast<Bytef *>(d); zs.avail_in = static_cast<uInt>(n);
        zs.next_out = out.data(); zs.avail_out = static_cast<uInt>(out.size());
        const int st = deflate(&zs, Z_FINISH);
        const size_t got = zs.total_out;
        const int endSt = deflateEnd(&zs);
        if (st != Z_STREAM_END) throw BamError("could not compress BGZF block");
        if (endSt != Z_OK) throw BamError("could not finalize BGZF compressor");
        out.resize(got);
    }
    // emitBlock (src/bgzf.cpp:250-283): appended to `dst`
    static void emitBlock(const uint8_t *d, size_t n, std::vector<uint8_t> &dst) {
        std::vector<uint8_t> comp;
        compressRaw(d, n, comp);
        const size_t total = 18 + comp.size() + 8;
        if (total > kMaxBlockSize) {
            if (n <= 1) throw BamError("BGZF block cannot fit compressed data");
            emitBlock(d, n / 2, dst);
            emitBlock(d + n / 2, n - n / 2, dst);
            return;
        }
        const size_t at = dst.size();
        dst.resize(at + total, 0);
        uint8_t *b = dst.data() + at;
        b[0] = 31; b[1] = 139; b[2] = 8; b[3] = 4; b[9] = 255;
        putU16(b + 10, 6); b[12] = 'B'; b[13] = 'C'; putU16(b + 14, 2); putU16(b + 16, static_cast<uint16_t>(total - 1));
        std::memcpy(b + 18, comp.data(), comp.size());
        putU32(b + total - 8, static_cast<uint32_t>(crc32(crc32(0L, Z_NULL, 0), d, static_cast<uInt>(n))));
        putU32(b + total - 4, static_cast<uint32_t>(n));
    }
    void flush(bool all) {
        const size_t full = pending_.size() / kMaxUncompressed, pieces = full + ((all && pending_.size() % kMaxUncompressed) ? 1 : 0);
        if (pieces == 0) return;
        std::vector<std::vector<uint8_t>> outs(pieces);
        std::vector<std::string> errs(pieces);
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < pieces;) {
                const size_t a = i * kMaxUncompressed, n = std::min(kMaxUncompressed, pending_.size() - a);
                try { emitBlock(pending_.data() + a, n, outs[i]); } catch (const std::exception &e) { errs[i] = e.what(); }
            }
        };
        tlsb::runThreads(std::min<size_t>(threads_, pieces), work);
        for (size_t i = 0; i < pieces; ++i) {
            if (!errs[i].empty()) throw BamError(errs[i]);
            if (fwrite(outs[i].data(), 1, outs[i].size(), f_) != outs[i].size()) throw BamError("failed while writing BAM output");
        }
        const size_t used = std::min(pending_.size(), pieces * kMaxUncompressed);
        pending_.erase(pending_.begin(), pending_.begin() + static_cast<std::ptrdiff_t>(used));
    }
};

// ---- BAM stream on top of the segments -----------------------------------------------------------------------------------
class BamStream {
public:
    BamStream(BgzfIn &in, size_t segmentBytes) : in_(in), seg_(segmentBytes) {}
    std::vector<uint8_t> buf;
    size_t off = 0;
    size_t avail() const { return buf.size() - off; }
    // BgzfReader::read: up to n bytes become available at buf[off...]; returns how many are
    size_t want(size_t n) {
        while (avail() < n) {
            if (off > (64u << 20) || (off && off == buf.size())) { buf.erase(buf.begin(), buf.begin() + static_cast<std::ptrdiff_t>(off)); off = 0; }
            if (!in_.nextSegment(buf, std::max(seg_, n - avail()))) break;
        }
        return std::min(n, avail());
    }
    void need(size_t n, const char *context) {               // readExact
        if (want(n) != n) throw BamError(std::string("truncated BAM ") + context);
    }
    // drops what has been consumed (the caller holds no pointers into buf any more)
    void compact() { if (off) { buf.erase(buf.begin(), buf.begin() + static_cast<std::ptrdiff_t>(off)); off = 0; } }
    // true when a new segment was appended; false at the end of the input.  Does not move what is in buf.
    bool more() { return in_.nextSegment(buf, seg_); }

private:
    BgzfIn &in_;
    size_t seg_;
};

struct RecRef { size_t raw, rawLen, seqOff; uint32_t seqLen; };   // offsets into BamStream::buf

// decodeSequence's checks (src/bam.cpp:122-163) on a complete record; fills where the packed SEQ sits
void locateSequence(const uint8_t *raw, size_t rawLen, RecRef &r) {
    const uint8_t *core = raw + 4;
    const size_t payloadSize = rawLen - 4;
    const size_t readNameLength = core[8], cigarCount = getU16(core + 12);
    const int32_t signedSeqLen = getI32(core + 16);
    if (readNameLength == 0 || signedSeqLen < 0) throw BamError("invalid BAM record lengths");
    const size_t seqLen = static_cast<size_t>(signedSeqLen);
    size_t so = kBamCoreSize;
    if (readNameLength > payloadSize - so) throw BamError("BAM read name exceeds block_size");
    so += readNameLength;
    const size_t cigarBytes = cigarCount * 4;
    if (cigarBytes > payloadSize - so) throw BamError("BAM CIGAR exceeds block_size");
    so += cigarBytes;
    const size_t packed = seqLen / 2 + seqLen % 2;
    if (packed > payloadSize - so) throw BamError("BAM sequence exceeds block_size");
    if (seqLen > payloadSize - (so + packed)) throw BamError("BAM record fields exceed block_size");
    if (core[kBamCoreSize + readNameLength - 1] != 0) throw BamError("BAM read name is not NUL-terminated");
    r.seqOff = 4 + so;
    r.seqLen = static_cast<uint32_t>(seqLen);
}

void copyHeader(BamStream &bs, BgzfOut &out) {                // copyBamHeader, src/bam.cpp:75-120
    auto pass = [&](size_t n, const char *context) {
        size_t left = n;
        while (left) {                                        // copyBytes: 64 KiB at a time, so that a huge header text does not need one buffer
            const size_t chunk = std::min<size_t>(left, 65536);
            bs.need(chunk, context);
            out.write(bs.buf.data() + bs.off, chunk);
            bs.off += chunk;
            left -= chunk;
        }
    };
    bs.need(4, "magic");
    if (std::memcmp(bs.buf.data() + bs.off, "BAM\1", 4) != 0) throw BamError("invalid BAM magic");
    pass(4, "magic");
    bs.need(4, "header length");
    const int32_t textLen = getI32(bs.buf.data() + bs.off);
    if (textLen < 0 || static_cast<size_t>(textLen) > kBamMaxHeaderText) throw BamError("invalid BAM header text length");
    pass(4, "header length");
    pass(static_cast<size_t>(textLen), "header text");
    bs.need(4, "reference count");
    const int32_t nRef = getI32(bs.buf.data() + bs.off);
    if (nRef < 0) throw BamError("invalid BAM reference count");
    pass(4, "reference count");
    for (int32_t i = 0; i < nRef; ++i) {
        bs.need(4, "reference name length");
        const int32_t nameLen = getI32(bs.buf.data() + bs.off);
        if (nameLen <= 0 || static_cast<size_t>(nameLen) > kBamMaxReferenceName) throw BamError("invalid BAM reference name length");
        pass(4, "reference name length");
        bs.need(static_cast<size_t>(nameLen), "reference name");
        if (bs.buf[bs.off + static_cast<size_t>(nameLen) - 1] != 0) throw BamError("BAM reference name is not NUL-terminated");
        pass(static_cast<size_t>(nameLen), "reference name");
        bs.need(4, "reference length");
        if (getI32(bs.buf.data() + bs.off) < 0) throw BamError("invalid BAM reference length");
        pass(4, "reference length");
    }
}

}  // namespace

namespace tlsb {

int runBamSubset(const Options &o, const tls_config &baseCfg) {
    StageTimer timer;
    tls_config cfg = baseCfg;
    cfg.min_block_len = o.minBlockLenSet ? o.minBlockLen : 42;          // src/read-filter.cpp:14-16
    cfg.outputs = 0;
    const unsigned threads = o.threads ? o.threads : std::max(1u, std::thread::hardware_concurrency());
    tls_ctx *ctx = nullptr;
    int ctxRc = TLS_OK;
    std::string ctxErr;
    std::thread ctxThread([&]() {                              // the CUDA context comes up while the input is inflated
        const char *devEnv = getenv("TLS_DEVICE");
        ctxRc = tls_create(&cfg, devEnv ? atoi(devEnv) : 0, &ctx);
        if (ctxRc != TLS_OK) ctxErr = tls_last_error();
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{ctxThread};

    std::string outputPath;
    FILE *out = stdout;
    uint64_t totalRecords = 0, passedRecords = 0, missingSeq = 0;
    bool missingEof = false;
    try {
        InputData in;
        if (o.inSequence.empty()) readAll(stdin, in);
        else if (!mapFile(o.inSequence, in)) throw BamError("cannot open BAM input '" + o.inSequence + "'");
        if (o.outRouteSet) {
            outputPath = o.outRoute + "/" + std::filesystem::path(o.inSequenceName).stem().string() + "_telomeric.bam";
            out = fopen(outputPath.c_str(), "wb");
            if (!out) throw BamError("cannot write telomeric records to '" + outputPath + "'");
        }
        timer.mark("open");
        const size_t segmentBytes = getenv("TLS_BAM_SEGMENT") ? strtoull(getenv("TLS_BAM_SEGMENT"), nullptr, 10) : (256ull << 20);
        BgzfIn bgzf(reinterpret_cast<const uint8_t *>(in.p), in.n, threads);
        BamStream bs(bgzf, segmentBytes);
        BgzfOut writer(out, threads);
        copyHeader(bs, writer);

        static const char bases[] = "=ACMGRSVTWYHKDBN";
        static uint16_t pairLut[256];
        for (int v = 0; v < 256; ++v) pairLut[v] = static_cast<uint16_t>(static_cast<uint8_t>(bases[v >> 4]) | (static_cast<uint8_t>(bases[v & 15]) << 8));
        std::vector<RecRef> recs;
        std::vector<uint8_t> seq;
        std::vector<uint64_t> offsets;
        std::vector<uint32_t> withSeq;                       // indices into recs of the records that go to the device
        double msInflate = 0, msDecode = 0, msGpu = 0, msOut = 0;
        auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
        const uint64_t batchBytes = 256ull << 20;

        // the records of recs[lo, hi) through the device, kept ones appended to the output in order
        auto scoreAndWrite = [&](size_t lo, size_t hi) {
            auto t0 = std::chrono::steady_clock::now();
            withSeq.clear(); offsets.assign(1, 0);
            for (size_t i = lo; i < hi; ++i)
                if (recs[i].seqLen) { withSeq.push_back(static_cast<uint32_t>(i)); offsets.push_back(offsets.back() + recs[i].seqLen); }
            seq.resize(offsets.back() + 256 + 2);
            std::atomic<size_t> next{0};
            auto decode = [&]() {
                for (size_t k; (k = next.fetch_add(64)) < withSeq.size();)
                    for (size_t j = k; j < std::min(withSeq.size(), k + 64); ++j) {
                        const RecRef &r = recs[withSeq[j]];
                        const uint8_t *src = bs.buf.data() + r.raw + r.seqOff;
                        uint8_t *dst = seq.data() + offsets[j];
                        const size_t pairs = r.seqLen / 2;
                        for (size_t q = 0; q < pairs; ++q) std::memcpy(dst + 2 * q, &pairLut[src[q]], 2);
                        if (r.seqLen & 1) dst[r.seqLen - 1] = static_cast<uint8_t>(bases[src[pairs] >> 4]);
                    }
            };
            runThreads(std::min<size_t>(threads, std::max<size_t>(1, withSeq.size() / 256)), decode);
            msDecode += since(t0);
            t0 = std::chrono::steady_clock::now();
            const uint8_t *pass = nullptr;
            tls_result res{};
            int ticket = -1;
            if (!withSeq.empty()) {
                if (ctxThread.joinable()) {
                    ctxThread.join();
                    if (ctxRc != TLS_OK) die("Error: tls_create failed: %s\n", ctxErr.c_str());
                }
                tls_batch tb{};
                tb.seq = seq.data(); tb.offsets = offsets.data(); tb.n_items = static_cast<uint32_t>(withSeq.size()); tb.mode = TLS_MODE_READS;
                check(tls_submit(ctx, &tb, &ticket), "tls_submit");
                check(tls_collect(ctx, ticket, &res), "tls_collect");
                pass = res.pass;
            }
            msGpu += since(t0);
            t0 = std::chrono::steady_clock::now();
            size_t j = 0;
            for (size_t i = lo; i < hi; ++i) {
                ++totalRecords;
                if (!recs[i].seqLen) { ++missingSeq; continue; }
                if (pass[j++]) { writer.write(bs.buf.data() + recs[i].raw, recs[i].rawLen); ++passedRecords; }
            }
            if (ticket >= 0) check(tls_release(ctx, ticket), "tls_release");
            msOut += since(t0);
        };

        for (bool done = false; !done;) {
            // ---- every complete record that is in the buffer (readRecord, src/bam.cpp:165-186) ----
            recs.clear();
            uint64_t seqBytes = 0;
            size_t lo = 0;
            for (;;) {
                if (bs.avail() < 4) {
                    auto t0 = std::chrono::steady_clock::now();
                    const bool got = recs.empty() ? (bs.compact(), bs.more()) : false;     // records hold offsets: refill only between rounds
                    msInflate += since(t0);
                    if (got) continue;
                    if (!recs.empty()) break;
                    if (bs.avail() == 0) { done = true; break; }
                    throw BamError("truncated BAM record size");
                }
                const int32_t blockSize = getI32(bs.buf.data() + bs.off);
                if (blockSize < static_cast<int32_t>(kBamCoreSize) || static_cast<size_t>(blockSize) > kBamMaxRecordSize)
                    throw BamError("invalid BAM record block_size");
                const size_t rawLen = static_cast<size_t>(blockSize) + 4;
                if (bs.avail() < rawLen) {
                    if (!recs.empty()) break;                 // finish this round first, then refill
                    auto t0 = std::chrono::steady_clock::now();
                    bs.compact();
                    bs.need(rawLen, "record");
                    msInflate += since(t0);
                }
                RecRef r{bs.off, rawLen, 0, 0};
                locateSequence(bs.buf.data() + bs.off, rawLen, r);
                bs.off += rawLen;
                recs.push_back(r);
                seqBytes += r.seqLen;
                if (seqBytes >= batchBytes) { scoreAndWrite(lo, recs.size()); lo = recs.size(); seqBytes = 0; }
            }
            if (lo < recs.size()) scoreAndWrite(lo, recs.size());
        }
        auto t0 = std::chrono::steady_clock::now();
        writer.finish();
        msOut += since(t0);
        missingEof = bgzf.physicalEof && !bgzf.sawEofBlock;
        timer.mark("subset");
        if (timer.on) fprintf(stderr, "[pipe]   of which: inflate %.1f ms, SEQ decode %.1f ms, device %.1f ms, deflate + write %.1f ms\n", msInflate, msDecode, msGpu, msOut);
    } catch (const std::exception &e) {
        if (out != stdout && out) {
            fclose(out);
            std::error_code ec;
            std::filesystem::remove(outputPath, ec);
        }
        fprintf(stderr, "Error: BAM subset failed: %s.\n", e.what());
        quit(EXIT_FAILURE);
    }
    if (out != stdout) fclose(out); else fflush(stdout);
    if (missingEof) fprintf(stderr, "Warning: BAM input is missing the BGZF EOF marker.\n");
    if (missingSeq > 0)
        fprintf(stderr, "BAM subset: skipped %llu record%s without SEQ.\n", static_cast<unsigned long long>(missingSeq), missingSeq == 1 ? "" : "s");
    fprintf(stderr, "BAM subset: kept %llu of %llu records.\n", static_cast<unsigned long long>(passedRecords), static_cast<unsigned long long>(totalRecords));
    if (!outputPath.empty()) fprintf(stderr, "Wrote telomeric records to %s.\n", outputPath.c_str());
    fflush(stderr);
    _exit(EXIT_SUCCESS);
}

}  // namespace tlsb
