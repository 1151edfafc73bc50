// teloscope-b200 — host pipeline around the C ABI for FASTA / FASTA.gz assemblies and FASTQ read subsets.
//
// What the reference does around scanSegment for a FASTA input, restated in C++17 on top of
// include/teloscope_b200.h so that the output files, stdout and stderr are byte-identical:
//   * flags of the scan (src/main.cpp:149-190 getopt table, :287-506 handlers, :590-674 messages);
//   * FASTA ingest with gfalibs' semantics (gfalibs/src/input-gfa.cpp:41-72, gfalibs/src/gfa.cpp:88-239):
//     header = text up to the first space, only '\n' leaves the sequence, maximal runs of N/n/X/x are gaps,
//     every other byte stays in its contig;
//   * Teloscope::walkPath (src/input.cpp:942-1041): every contig of a path is one scanSegment call with its
//     absPos — here: one item of a batch that goes to the GPU through tls_submit / tls_collect, two batches
//     in flight;
//   * labelTerminalBlocks (src/teloscope.cpp:259-383), the writers (:661-859), file set (:886-920), the path
//     and assembly summaries (:950-1055), getStats (src/tools.cpp:22-51), computeN50 (include/teloscope.h:224).
//   * --fastq-subset (Input::readFastqSubset, src/input.cpp:735-833): FASTQ / FASTQ.gz records whose read has a
//     Teloscope-valid telomeric block, written back verbatim in input order (stdout, or a file with -o).
// Not here (the reference's other front ends; this binary says so and exits with status 2): GFA input,
// --bam-subset, the record filters (--include-bed ...), --plot-report, stdin.
//
// The float tracks are computed on the host from the device's integer window counters with the reference's
// own expressions (include/teloscope.h:199-214, src/teloscope.cpp:788-803): compiled with -ffp-contract=off.
#include <emmintrin.h>
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdarg>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <sstream>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pipeline.hpp"

using namespace tlsb;

namespace {

std::string upper(std::string s) {
    for (char &c : s) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    return s;
}
std::string revCom(const std::string &s) {          // gfalibs revCom: complement of ACGT (either case), other bytes kept
    std::string r(s.rbegin(), s.rend());
    for (char &c : r) {
        switch (c) {
            case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
            case 'a': c = 't'; break; case 'c': c = 'g'; break; case 'g': c = 'c'; break; case 't': c = 'a'; break;
            default: break;
        }
    }
    return r;
}
bool hasDigit(const std::string &s) { return std::any_of(s.begin(), s.end(), [](unsigned char c) { return std::isdigit(c); }); }

int parseInt(const char *arg, const char *invalidMsg, const char *rangeMsg) {      // std::stoi + "> 0" of the reference's handlers
    try {
        const int v = std::stoi(arg);
        if (v <= 0) die("%s", rangeMsg);
        return v;
    } catch (const std::exception &) {
        die(invalidMsg, arg);
    }
}

void setInputFile(Options &o, const char *path) {                                    // src/main.cpp:92-101
    std::error_code ec;
    if (!std::filesystem::exists(path, ec)) die("Error: Input file '%s' does not exist.\n", path);
    const std::filesystem::path real = std::filesystem::canonical(path);
    o.inSequence = real.string();
    o.inSequencePrefix = real.parent_path().string();
    o.inSequenceName = real.filename().string();
    if (!o.outRouteSet) o.outRoute = o.inSequencePrefix;
}

Options parseArgs(int argc, char **argv) {
    Options o;
    static int cmdFlag = 0, verboseFlag = 0;
    static struct option longOptions[] = {
        {"input-sequence", required_argument, 0, 'f'}, {"output", required_argument, 0, 'o'},
        {"include-bed", required_argument, 0, 0}, {"exclude-bed", required_argument, 0, 0},
        {"include-prefix", required_argument, 0, 0}, {"exclude-prefix", required_argument, 0, 0},
        {"patterns", required_argument, 0, 'p'}, {"window", required_argument, 0, 'w'}, {"step", required_argument, 0, 's'},
        {"canonical", required_argument, 0, 'c'}, {"threads", required_argument, 0, 'j'},
        {"terminal-limit", required_argument, 0, 't'}, {"max-match-distance", required_argument, 0, 'k'},
        {"max-block-distance", required_argument, 0, 'd'}, {"min-block-length", required_argument, 0, 'l'},
        {"min-block-density", required_argument, 0, 'y'}, {"edit-distance", required_argument, 0, 'x'},
        {"out-fasta", no_argument, 0, 'a'}, {"out-win-repeats", no_argument, 0, 'r'}, {"out-gc", no_argument, 0, 'g'},
        {"out-entropy", no_argument, 0, 'e'}, {"out-matches", no_argument, 0, 'm'}, {"out-its", no_argument, 0, 'i'},
        {"ultra-fast", no_argument, 0, 'u'}, {"manual-curation", no_argument, 0, 'n'}, {"plot-report", no_argument, 0, 0},
        {"fastq-subset", no_argument, 0, 0}, {"bam-subset", no_argument, 0, 0}, {"gpus", required_argument, 0, 0}, {"verbose", no_argument, &verboseFlag, 1},
        {"cmd", no_argument, &cmdFlag, 1}, {"version", no_argument, 0, 'v'}, {"help", no_argument, 0, 'h'}, {0, 0, 0, 0}};
    for (;;) {
        int idx = 0;
        const int c = getopt_long(argc, argv, "-:f:j:o:p:s:w:c:t:k:d:l:y:x:argemivhun", longOptions, &idx);
        if (c == -1) break;
        switch (c) {
            case ':':
                if (optopt == 0 && optind > 0 && optind <= argc && strncmp(argv[optind - 1], "--", 2) == 0)
                    fprintf(stderr, "Error: Option %s is missing a required argument\n", argv[optind - 1]);
                else
                    fprintf(stderr, "Error: Option -%c is missing a required argument\n", optopt);
                exit(EXIT_FAILURE);
            case 0: {
                const char *name = longOptions[idx].name;
                if (longOptions[idx].flag) break;                      // --verbose / --cmd
                if (!strcmp(name, "plot-report")) unsupported("--plot-report");
                if (!strcmp(name, "fastq-subset")) { o.fastqSubset = true; break; }
                if (!strcmp(name, "gpus")) {               // not a reference flag: devices the scan of one job is spread over
                    o.gpus = atoi(optarg);
                    if (o.gpus < 1 || o.gpus > 64) die("Error: --gpus must be between 1 and 64.\n");
                    break;
                }
                if (!strcmp(name, "bam-subset")) { o.bamSubset = true; break; }
                unsupported("record filtering (--include-bed / --exclude-bed / --include-prefix / --exclude-prefix)");
            }
            case 1:
                if (o.inSequence.empty()) setInputFile(o, optarg);
                else fprintf(stderr, "Warning: Ignoring extra positional argument '%s'\n", optarg);
                break;
            case 'f': setInputFile(o, optarg); break;
            case 'o':
                o.outRoute = optarg;
                o.outRouteSet = true;
                if (o.outRoute.empty()) exit(EXIT_FAILURE);
                try {
                    if (!std::filesystem::exists(o.outRoute)) std::filesystem::create_directories(o.outRoute);
                } catch (const std::filesystem::filesystem_error &e) {
                    die("Error: Cannot create output directory '%s': %s\n", o.outRoute.c_str(), e.code().message().c_str());
                }
                break;
            case 'j': o.threads = static_cast<unsigned>(std::max(0, atoi(optarg))); break;   // host threads (parse, formatting); the scan runs on the GPU
            case 'c': {
                if (!optarg || !*optarg) {
                    fprintf(stderr, "Warning: Empty canonical pattern provided, using default vertebrate TTAGGG.\n");
                    break;
                }
                const std::string can = upper(optarg);
                if (hasDigit(can)) die("Error: Canonical pattern '%s' contains numerical characters.\n", can.c_str());
                const std::string rc = revCom(can);
                if (can <= rc) { o.canonicalFwd = can; o.canonicalRev = rc; }
                else { o.canonicalFwd = rc; o.canonicalRev = can; }
                fprintf(stderr, "Setting canonical pattern: %s and its reverse complement: %s\n", o.canonicalFwd.c_str(),
                        o.canonicalRev.c_str());
                break;
            }
            case 'p': {
                o.hasInputPatterns = true;
                if (!optarg || !*optarg) {
                    fprintf(stderr, "Warning: Empty pattern list provided, using default: TTAGGG, CCCTAA\n");
                    o.rawPatterns = {"TTAGGG", "CCCTAA"};
                    break;
                }
                o.rawPatterns.clear();
                std::istringstream ps(optarg);
                std::string pat;
                while (std::getline(ps, pat, ',')) {
                    if (pat.empty()) continue;
                    if (hasDigit(pat)) die("Error: Pattern '%s' contains numerical characters.\n", pat.c_str());
                    o.rawPatterns.push_back(upper(pat));
                }
                break;
            }
            case 'w':
                o.windowSize = parseInt(optarg, "Error: Invalid window size '%s'. Must be a number.\n",
                                        "Error: Window size (-w or --window) must be > 0.\n");
                break;
            case 's':
                o.step = parseInt(optarg, "Error: Invalid step size '%s'. Must be a number.\n",
                                  "Error: Step size (-s or --step) must be > 0.\n");
                break;
            case 't':
                o.terminalLimit = parseInt(optarg, "Error: Invalid terminal limit '%s'. Must be a number.\n",
                                           "Error: Terminal limit (-t or --terminal-limit) must be > 0.\n");
                break;
            case 'k':
                o.maxMatchDist = static_cast<unsigned short>(parseInt(optarg, "Error: Invalid max match distance '%s'. Must be a number.\n",
                                                                        "Error: Max match distance (-k/--max-match-distance) must be > 0.\n"));
                break;
            case 'l':
                o.minBlockLen = static_cast<unsigned short>(parseInt(optarg, "Error: Invalid min block length '%s'. Must be a number.\n",
                                                                       "Error: Min block length (-l or --min-block-length) must be > 0.\n"));
                o.minBlockLenSet = true;
                break;
            case 'd':
                o.maxBlockDist = static_cast<unsigned short>(parseInt(optarg, "Error: Invalid max block distance '%s'. Must be a number.\n",
                                                                        "Error: Max block distance (-d or --max-block-distance) must be > 0.\n"));
                break;
            case 'y':
                try {
                    const float v = std::stof(optarg);
                    if (v < 0.0f || v > 1.0f) die("Error: Min block density (-y/--min-block-density) must be in the range [0,1].\n");
                    o.minBlockDensity = v;
                } catch (const std::exception &) {
                    die("Error: Invalid min block density '%s'. Must be a number [0,1].\n", optarg);
                }
                break;
            case 'x':
                try {
                    const int v = std::stoi(optarg);
                    if (v < 0 || v > 2) die("Error: Edit distance (-x/--edit-distance) must be in the range [0,2].\n");
                    o.editDistance = static_cast<uint8_t>(v);
                } catch (const std::exception &) {
                    die("Error: Invalid edit distance '%s'. Must be a number [0,2].\n", optarg);
                }
                break;
            case 'a': o.outFasta = true; o.ultraFastMode = false; break;
            case 'r': o.outWinRepeats = true; o.ultraFastMode = false; break;
            case 'g': o.outGC = true; o.ultraFastMode = false; break;
            case 'e': o.outEntropy = true; o.ultraFastMode = false; break;
            case 'm': o.outMatches = true; o.ultraFastMode = false; break;
            case 'i': o.outITS = true; o.ultraFastMode = false; break;
            case 'u':
                if (o.outWinRepeats || o.outGC || o.outEntropy || o.outITS || o.outMatches) {
                    o.ultraFastMode = false;
                    fprintf(stderr, "Ignoring -u: -r/-g/-e/-i/-m request genome-wide scanning.\n");
                } else {
                    o.ultraFastMode = true;
                    fprintf(stderr, "Fast mode: Only scanning terminal regions.\n");
                }
                break;
            case 'n': o.manualCuration = true; break;
            case 'v': printf("/// teloscope-b200 (B200 FASTA pipeline of Teloscope's scan path)\n"); exit(0);
            case 'h':
                printf("teloscope-b200 input.[fa|fa.gz] [options]\n"
                       "Scan flags as in teloscope: -f -o -c -p -x -w -s -t -k -d -l -y -r -g -e -m -i -u -n -j (ignored).\n");
                exit(0);
            default:
                fprintf(stderr, "Error: Unknown option.\n");
                exit(EXIT_FAILURE);
        }
    }
    o.cmdFlag = cmdFlag;
    return o;
}

// ---- FASTA ingest -------------------------------------------------------------------------------------
struct Gap { uint64_t start; uint32_t length; };
struct PathRec {
    uint32_t seqPos = 0;
    std::string header;
    uint64_t pathSize = 0;
    std::vector<Gap> gaps;
    std::vector<tls_block> terminal, its;
    uint64_t nCanonical = 0, nWindows = 0;
    std::string terminalLabel;
    int scaffoldType = 6;                       // index into kTypeNames: none
    std::vector<uint8_t> isLongest;
};
struct Item { uint32_t path; uint64_t absPos, offset; uint32_t len; };

// bit i of the result: byte i of the 16 at p is one of N n X x (gap) / is '\n'
inline void classify16(const char *p, unsigned &gap, unsigned &nl) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p));
    const __m128i u = _mm_or_si128(v, _mm_set1_epi8(0x20));                       // fold case: n / x
    gap = static_cast<unsigned>(_mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(u, _mm_set1_epi8('n')), _mm_cmpeq_epi8(u, _mm_set1_epi8('x')))));
    nl = static_cast<unsigned>(_mm_movemask_epi8(_mm_cmpeq_epi8(v, _mm_set1_epi8('\n'))));
}

struct RecordSpan { size_t begin, end, bodyBegin; };          // [begin, end): header line + body, without the leading '>'

// pass 1 over a piece of a record body: contig bases (everything but '\n' and gap characters) and path positions
// (everything but '\n')
void countPiece(const char *p, const char *e, uint64_t &bases, uint64_t &positions) {
    uint64_t skipped = 0, newlines = 0;
    const char *q = p;
    for (; q + 16 <= e; q += 16) {
        unsigned gap, nl;
        classify16(q, gap, nl);
        skipped += static_cast<uint64_t>(__builtin_popcount(gap | nl));
        newlines += static_cast<uint64_t>(__builtin_popcount(nl));
    }
    for (; q < e; ++q) { skipped += (*q == '\n' || *q == 'N' || *q == 'n' || *q == 'X' || *q == 'x'); newlines += *q == '\n'; }
    bases = static_cast<uint64_t>(e - p) - skipped;
    positions = static_cast<uint64_t>(e - p) - newlines;
}

// pass 2: the record's contigs copied back to back to out (newlines and gaps dropped), its items and gaps listed
void parseRecord(const char *p, const char *e, uint8_t *out, uint64_t outCap, uint64_t outOffset, uint32_t pathIdx, std::vector<Item> &items,
                 std::vector<Gap> &gaps, uint64_t &pathSize) {
    uint64_t pos = 0, runStart = 0, written = 0, segOffset = outOffset;
    int state = -1;                                           // -1 nothing yet, 0 in a contig, 1 in a gap
    auto closeRun = [&]() {
        if (state == 1) gaps.push_back({runStart, static_cast<uint32_t>(pos - runStart)});
        else if (state == 0) items.push_back({pathIdx, runStart, segOffset, static_cast<uint32_t>(pos - runStart)});
    };
    auto byteStep = [&](uint8_t c) {
        if (c == '\n') return;
        const int st = (c == 'N' || c == 'n' || c == 'X' || c == 'x') ? 1 : 0;
        if (st != state) {
            closeRun();
            state = st; runStart = pos; segOffset = outOffset + written;
        }
        if (st == 0) out[written++] = c;
        ++pos;
    };
    const char *q = p;
    while (q < e) {
        if (state == 0 && q + 16 <= e) {                      // fast path inside a contig: whole 16-byte pieces without gap bytes
            unsigned gap, nl;
            classify16(q, gap, nl);
            if (gap == 0) {
                if (nl == 0) {
                    memcpy(out + written, q, 16);
                    written += 16; pos += 16;
                } else if ((nl & (nl - 1)) == 0 && q + 33 <= e && written + 32 <= outCap) {
                    // exactly one newline (the usual line end): the bytes before it, then the 15 after it, as two
                    // overlapping 16-byte copies (the second one overwrites the newline's slot and what follows)
                    const int k = __builtin_ctz(nl);
                    memcpy(out + written, q, 16);
                    memcpy(out + written + k, q + k + 1, 16);
                    written += 15; pos += 15;
                } else {
                    for (int i = 0; i < 16; ++i)
                        if (!((nl >> i) & 1u)) { out[written++] = static_cast<uint8_t>(q[i]); ++pos; }
                }
                q += 16;
                continue;
            }
        }
        if (state == 1 && q + 16 <= e) {                      // fast path inside a gap: 16 bytes of gap characters / newlines
            unsigned gap, nl;
            classify16(q, gap, nl);
            if ((gap | nl) == 0xFFFFu) {
                pos += static_cast<uint64_t>(__builtin_popcount(gap));
                q += 16;
                continue;
            }
        }
        byteStep(static_cast<uint8_t>(*q++));
    }
    closeRun();
    pathSize = pos;
}

void parseFasta(const InputData &in, unsigned threads, std::vector<PathRec> &paths, std::vector<Item> &items,
                uint8_t *&seq, uint64_t &seqBytes) {
    if (in.n == 0 || in.p[0] != '>') die("Error: '%s' does not look like a FASTA file.\n", "input");
    // records: a record runs to the next '>' wherever it stands (gfalibs reads the sequence with getline(..., '>'))
    // the '>' bytes are looked for in parallel (a serial memchr over a 3 GB file costs 0.3 s)
    std::vector<RecordSpan> recs;
    {
        const size_t nChunks = std::max<size_t>(1, std::min<size_t>(threads * 4, in.n / (4u << 20) + 1)), chunk = (in.n + nChunks - 1) / nChunks;
        std::vector<std::vector<size_t>> found(nChunks);
        parallelFor(nChunks, threads, [&](size_t c) {
            const size_t lo = std::max<size_t>(1, c * chunk), hi = std::min(in.n, (c + 1) * chunk);
            for (size_t q = lo; q < hi;) {
                const char *g = static_cast<const char *>(memchr(in.p + q, '>', hi - q));
                if (!g) break;
                found[c].push_back(static_cast<size_t>(g - in.p));
                q = static_cast<size_t>(g - in.p) + 1;
            }
        });
        size_t p = 1;
        auto close = [&](size_t end) {
            const char *nlp = static_cast<const char *>(memchr(in.p + p, '\n', end - p));
            recs.push_back({p, end, nlp ? static_cast<size_t>(nlp - in.p) + 1 : end});
            p = end + 1;
        };
        for (const auto &v : found) for (size_t g : v) close(g);
        close(in.n);
    }
    paths.resize(recs.size());
    // A record is cut into pieces of kPiece bytes that are counted (pass 1) and copied (pass 2) in parallel: the largest
    // record (a chromosome) does not bound the wall time.  Pieces start in an unknown state; runs of contig bases / gap
    // characters that continue across a cut are stitched afterwards.
    const size_t kPiece = getenv("TLS_PARSE_PIECE") ? std::max<size_t>(64, strtoull(getenv("TLS_PARSE_PIECE"), nullptr, 10)) : (8u << 20);   // tests shrink it
    struct Piece { size_t rec; const char *p, *e; uint64_t bases = 0, positions = 0, outOffset = 0, posBase = 0; std::vector<Item> items; std::vector<Gap> gaps; };
    std::vector<Piece> pieces;
    std::vector<size_t> firstPiece(recs.size() + 1, 0);
    for (size_t r = 0; r < recs.size(); ++r) {
        firstPiece[r] = pieces.size();
        const char *p = in.p + recs[r].bodyBegin, *e = in.p + recs[r].end;
        if (p >= e) { Piece pc; pc.rec = r; pc.p = pc.e = e; pieces.push_back(std::move(pc)); continue; }
        for (; p < e; p += kPiece) { Piece pc; pc.rec = r; pc.p = p; pc.e = std::min(e, p + kPiece); pieces.push_back(std::move(pc)); }
    }
    firstPiece[recs.size()] = pieces.size();
    parallelFor(pieces.size(), threads, [&](size_t k) { countPiece(pieces[k].p, pieces[k].e, pieces[k].bases, pieces[k].positions); });
    uint64_t outAcc = 0;
    for (size_t r = 0; r < recs.size(); ++r) {
        uint64_t posAcc = 0;
        for (size_t k = firstPiece[r]; k < firstPiece[r + 1]; ++k) {
            pieces[k].outOffset = outAcc; pieces[k].posBase = posAcc;
            outAcc += pieces[k].bases; posAcc += pieces[k].positions;
        }
        paths[r].pathSize = posAcc;
    }
    seqBytes = outAcc;
    {   // the contig buffer: 2 MiB aligned and advised for huge pages (3 GB of first-touch faults otherwise); it lives until exit
        const size_t cap = (static_cast<size_t>(seqBytes) + 256 + (2u << 20) - 1) & ~static_cast<size_t>((2u << 20) - 1);
        void *mem = nullptr;
        if (posix_memalign(&mem, 2u << 20, cap) != 0 || !mem) die("Error: out of memory (%zu bytes for the sequence).\n", cap);
        madvise(mem, cap, MADV_HUGEPAGE);
        seq = static_cast<uint8_t *>(mem);
    }
    memset(seq + seqBytes, 0, 256);
    parallelFor(pieces.size(), threads, [&](size_t k) {
        Piece &pc = pieces[k];
        uint64_t ignored = 0;
        parseRecord(pc.p, pc.e, seq + pc.outOffset, pc.bases, pc.outOffset, static_cast<uint32_t>(pc.rec), pc.items, pc.gaps, ignored);
    });
    std::vector<std::vector<Item>> recItems(recs.size());
    parallelFor(recs.size(), threads, [&](size_t r) {
        PathRec &rec = paths[r];
        rec.seqPos = static_cast<uint32_t>(r);
        const size_t headEnd = recs[r].bodyBegin > recs[r].begin && recs[r].bodyBegin <= recs[r].end && in.p[recs[r].bodyBegin - 1] == '\n'
                                   ? recs[r].bodyBegin - 1 : recs[r].end;
        const char *sp = static_cast<const char *>(memchr(in.p + recs[r].begin, ' ', headEnd - recs[r].begin));
        rec.header.assign(in.p + recs[r].begin, sp ? sp : in.p + headEnd);
        rec.header.erase(std::remove(rec.header.begin(), rec.header.end(), '\r'), rec.header.end());   // walkPath, src/input.cpp:950
        // stitch the pieces: positions become path positions; a run that reaches the end of a piece and one of the same
        // kind that starts the next piece are one run
        std::vector<Item> &its = recItems[r];
        for (size_t k = firstPiece[r]; k < firstPiece[r + 1]; ++k) {
            Piece &pc = pieces[k];
            for (Item &it : pc.items) {
                it.absPos += pc.posBase;
                if (!its.empty() && it.absPos == pc.posBase && its.back().absPos + its.back().len == it.absPos &&
                    its.back().offset + its.back().len == it.offset) {
                    if (static_cast<uint64_t>(its.back().len) + it.len > 0xFFFFFFFFull) die("Error: a contig is longer than 2^32-1 bases.\n");
                    its.back().len += it.len;
                } else its.push_back(it);
            }
            for (Gap &g : pc.gaps) {
                g.start += pc.posBase;
                if (!rec.gaps.empty() && g.start == pc.posBase && rec.gaps.back().start + rec.gaps.back().length == g.start) rec.gaps.back().length += g.length;
                else rec.gaps.push_back(g);
            }
            std::vector<Item>().swap(pc.items);
        }
    });
    size_t total = 0;
    for (auto &v : recItems) total += v.size();
    items.reserve(total);
    for (auto &v : recItems) items.insert(items.end(), v.begin(), v.end());
}

// ---- output ---------------------------------------------------------------------------------------------
struct OutFile {
    std::string path, buf;
    FILE *fh = nullptr;
    bool open(const std::string &p) {
        path = p;
        fh = fopen(p.c_str(), "wb");
        if (!fh) die("Error: Could not open '%s' for writing.\n", p.c_str());
        buf.reserve(1 << 22);
        return true;
    }
    void flushIfBig() { if (buf.size() > (1u << 22) - 4096) flush(); }
    void flush() { if (fh && !buf.empty()) { fwrite(buf.data(), 1, buf.size(), fh); buf.clear(); } }
    void close() { flush(); if (fh) { fclose(fh); fh = nullptr; } }
};
inline void appendU64(std::string &s, uint64_t v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = static_cast<char>('0' + v % 10); v /= 10; } while (v);
    while (n) s.push_back(tmp[--n]);
}
inline void appendFloat(std::string &s, float v) {              // std::ostream << float, default precision
    // the window tracks take few distinct values (ratios of small counts): a direct-mapped per-thread cache of the
    // printed text, keyed by the float's bits, saves most of the snprintf calls
    struct Entry { uint32_t bits; uint8_t len; char txt[19]; };
    thread_local Entry cache[4096] = {};
    uint32_t bits;
    memcpy(&bits, &v, 4);
    Entry &e = cache[(bits * 2654435761u) >> 20];
    if (e.len == 0 || e.bits != bits) {
        char tmp[40];
        const int n = snprintf(tmp, sizeof tmp, "%g", static_cast<double>(v));
        if (n <= 0 || n > 19) { s.append(tmp, static_cast<size_t>(std::max(n, 0))); return; }
        e.bits = bits; e.len = static_cast<uint8_t>(n);
        memcpy(e.txt, tmp, static_cast<size_t>(n));
    }
    s.append(e.txt, e.len);
}
inline void windowLine(std::string &s, const std::string &header, uint64_t a, uint64_t b, float v) {
    s += header; s.push_back('\t'); appendU64(s, a); s.push_back('\t'); appendU64(s, b); s.push_back('\t'); appendFloat(s, v);
    s.push_back('\n');
}
float shannonEntropy(const uint32_t nt[4], uint32_t windowSize) {         // include/teloscope.h:199-208
    float entropy = 0.0;
    for (int i = 0; i < 4; ++i) {
        if (nt[i] > 0) {
            const float probability = static_cast<float>(nt[i]) / windowSize;
            entropy -= probability * std::log2(probability);
        }
    }
    return std::round(entropy * 1000.0f) / 1000.0f;
}
float gcContent(const uint32_t nt[4], uint32_t windowSize) {               // include/teloscope.h:211-214
    const uint32_t gcCount = nt[1] + nt[2];
    return static_cast<float>(gcCount) / windowSize * 100.0;
}

const char *kTypeNames[] = {"t2t", "gapped_t2t", "misassembly", "gapped_misassembly", "incomplete", "gapped_incomplete",
                            "none", "gapped_none", "discordant", "gapped_discordant"};
enum { T2T = 0, MISASSEMBLY = 2, INCOMPLETE = 4, NONE = 6, DISCORDANT = 8 };

void labelTerminalBlocks(PathRec &p, uint32_t terminalLimit) {             // src/teloscope.cpp:259-383
    auto &blocks = p.terminal;
    const int gapped = p.gaps.empty() ? 0 : 1;
    p.isLongest.assign(blocks.size(), 0);
    p.terminalLabel.clear();
    if (blocks.empty()) { p.scaffoldType = NONE + gapped; return; }
    std::sort(blocks.begin(), blocks.end(), [](const tls_block &a, const tls_block &b) { return a.start < b.start; });
    std::vector<size_t> scaffold;
    for (size_t i = 0; i < blocks.size(); ++i) {
        const uint64_t blockEnd = blocks[i].start + blocks[i].len;
        if (blocks[i].start < terminalLimit || blockEnd > p.pathSize - terminalLimit) scaffold.push_back(i);
    }
    for (const auto &b : blocks) {
        p.terminalLabel += b.label;
        if (!b.has_valid_or) p.terminalLabel += '*';
    }
    long longestP = -1, longestQ = -1;
    uint64_t maxP = 0, maxQ = 0;
    for (size_t i : scaffold) {
        if (blocks[i].label == 'p' && blocks[i].can_cov > maxP) { longestP = static_cast<long>(i); maxP = blocks[i].can_cov; }
        else if (blocks[i].label == 'q' && blocks[i].can_cov > maxQ) { longestQ = static_cast<long>(i); maxQ = blocks[i].can_cov; }
    }
    if (longestP >= 0) p.isLongest[longestP] = 1;
    if (longestQ >= 0) p.isLongest[longestQ] = 1;
    for (size_t i = 0, j = 0; i < blocks.size(); ++i) {
        if (p.isLongest[i]) p.terminalLabel[j] = static_cast<char>(std::toupper(static_cast<unsigned char>(p.terminalLabel[j])));
        ++j;
        if (j < p.terminalLabel.size() && p.terminalLabel[j] == '*') ++j;
    }
    const bool hasP = longestP >= 0, hasQ = longestQ >= 0;
    if ((hasP && !blocks[longestP].has_valid_or) || (hasQ && !blocks[longestQ].has_valid_or)) { p.scaffoldType = DISCORDANT + gapped; return; }
    if (hasP && hasQ) {
        p.scaffoldType = (blocks[longestP].start < blocks[longestQ].start ? T2T : MISASSEMBLY) + gapped;
        return;
    }
    for (int arm = 0; arm < 2; ++arm) {
        const long longest = arm == 0 ? longestP : longestQ;
        if (longest < 0) continue;
        const char lab = arm == 0 ? 'p' : 'q';
        for (size_t i : scaffold)
            if (blocks[i].label == lab && static_cast<long>(i) != longest && blocks[i].has_valid_or) { p.scaffoldType = MISASSEMBLY + gapped; return; }
    }
    if (!hasP && !hasQ) { p.scaffoldType = NONE + gapped; return; }
    p.scaffoldType = INCOMPLETE + gapped;
}

uint64_t computeN50(std::vector<uint64_t> &lengths) {                       // include/teloscope.h:224-236
    if (lengths.empty()) return 0;
    std::sort(lengths.begin(), lengths.end(), [](uint64_t a, uint64_t b) { return a > b; });
    uint64_t total = 0;
    for (uint64_t l : lengths) total += l;
    uint64_t cum = 0;
    for (uint64_t l : lengths) {
        cum += l;
        if (cum * 2 >= total) return l;
    }
    return lengths.back();
}

void blockRow(std::string &s, const std::string &header, const tls_block &b, uint64_t pathSize) {
    s += header; s.push_back('\t'); appendU64(s, b.start); s.push_back('\t'); appendU64(s, b.start + b.len); s.push_back('\t');
    appendU64(s, b.len); s.push_back('\t'); s.push_back(b.label); s.push_back('\t'); appendU64(s, b.fwd_count); s.push_back('\t');
    appendU64(s, b.rev_count); s.push_back('\t'); appendU64(s, b.can_count); s.push_back('\t'); appendU64(s, b.noncan_count);
    s.push_back('\t'); appendU64(s, pathSize);
}

// ---- --fastq-subset (Input::readFastqSubset, src/input.cpp:735-833; readFastqRecord :113-138) -----------------------------
// Reads are taken four lines at a time (blank lines before a header are skipped), their sequence lines go to the
// GPU as TLS_MODE_READS batches (ReadTelomereFilter::matches: trailing '\r' dropped, whole read terminal, -l 42 unless
// given), and the records whose read passes are written back exactly as they were read, in input order.
[[noreturn]] void fastqError(uint64_t recordNumber, const char *message) {
    if (recordNumber > 0) fprintf(stderr, "Error: FASTQ record %llu: %s.\n", static_cast<unsigned long long>(recordNumber), message);
    else fprintf(stderr, "Error: %s.\n", message);
    quit(EXIT_FAILURE);
}

int runFastqSubset(const Options &o, const tls_config &baseCfg) {
    StageTimer timer;
    tls_config cfg = baseCfg;
    cfg.min_block_len = o.minBlockLenSet ? o.minBlockLen : 42;          // src/read-filter.cpp:14-16
    cfg.outputs = 0;
    tls_ctx *ctx = nullptr;
    int ctxRc = TLS_OK;
    std::string ctxErr;
    std::thread ctxThread([&]() {                              // the CUDA context comes up while the input is mapped / inflated
        const char *devEnv = getenv("TLS_DEVICE");
        ctxRc = tls_create(&cfg, devEnv ? atoi(devEnv) : 0, &ctx);
        if (ctxRc != TLS_OK) ctxErr = tls_last_error();
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{ctxThread};
    InputData in;
    if (o.inSequence.empty()) {
        readAll(stdin, in);
        if (in.n > 1 && static_cast<unsigned char>(in.p[0]) == 0x1f) {          // src/main.cpp:576-583
            fprintf(stderr, "Error: Compressed input on stdin is not supported. Decompress first:\n");
            fprintf(stderr, "  zcat file.fa.gz | teloscope -o results/\n");
            quit(EXIT_FAILURE);
        }
    } else {
        loadInput(o.inSequence, in);
    }
    timer.mark("read file");
    if (in.n == 0) fastqError(0, "FASTQ input is empty");
    if (in.p[0] != '@') fastqError(0, "FASTQ input must start with '@'");
    FILE *out = stdout;
    std::string outPath;
    if (o.outRouteSet) {                                      // src/main.cpp:699-713
        outPath = o.outRoute + "/" + o.inSequenceName + "_telomeric.fastq";
        out = fopen(outPath.c_str(), "wb");
        if (!out) die("Error: Cannot write telomeric reads to '%s'.\n", outPath.c_str());
    }
    struct Line { size_t b, e; };
    struct Rec { Line header, sequence, plus, quality; };
    struct Batch { std::vector<Rec> recs; std::vector<uint8_t> seq; std::vector<uint64_t> offsets; int ticket = -1; };
    size_t pos = 0;
    auto nextLine = [&](Line &l) {                            // std::getline
        if (pos >= in.n) return false;
        const char *nl = static_cast<const char *>(memchr(in.p + pos, '\n', in.n - pos));
        l.b = pos;
        l.e = nl ? static_cast<size_t>(nl - in.p) : in.n;
        pos = nl ? l.e + 1 : in.n;
        return true;
    };
    auto logicalLen = [&](const Line &l) { return l.e - l.b - ((l.e > l.b && in.p[l.e - 1] == '\r') ? 1 : 0); };
    uint64_t recordNumber = 0, totalReads = 0, passedReads = 0;
    const uint64_t batchBytes = getenv("TLS_BATCH_BYTES") ? strtoull(getenv("TLS_BATCH_BYTES"), nullptr, 10) : (256ull << 20);
    std::vector<std::unique_ptr<Batch>> flight, pool;
    std::string outBuf;
    auto consume = [&](Batch &b) {
        tls_result r{};
        check(tls_collect(ctx, b.ticket, &r), "tls_collect");
        for (size_t j = 0; j < b.recs.size(); ++j) {
            ++totalReads;
            if (!r.pass[j]) continue;
            ++passedReads;
            for (const Line *l : {&b.recs[j].header, &b.recs[j].sequence, &b.recs[j].plus, &b.recs[j].quality}) {
                outBuf.append(in.p + l->b, l->e - l->b);
                outBuf.push_back('\n');
            }
            if (outBuf.size() > (1u << 22)) { fwrite(outBuf.data(), 1, outBuf.size(), out); outBuf.clear(); }
        }
        check(tls_release(ctx, b.ticket), "tls_release");
    };
    bool more = true;
    double msPack = 0, msSubmit = 0, msCollect = 0;
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    while (more || !flight.empty()) {
        if (more && flight.size() < 2) {
            auto t0 = std::chrono::steady_clock::now();
            std::unique_ptr<Batch> b;
            if (!pool.empty()) { b = std::move(pool.back()); pool.pop_back(); }     // buffers keep their (touched) capacity
            else { b = std::make_unique<Batch>(); b->seq.reserve(batchBytes + (64u << 20)); }
            b->recs.clear(); b->seq.clear(); b->offsets.clear();
            b->offsets.push_back(0);
            while (b->seq.size() < batchBytes && b->recs.size() < (1u << 22)) {
                Rec rec;
                do {
                    if (!nextLine(rec.header)) { more = false; break; }
                } while (logicalLen(rec.header) == 0);
                if (!more) break;
                if (!nextLine(rec.sequence) || !nextLine(rec.plus) || !nextLine(rec.quality)) fastqError(recordNumber + 1, "truncated FASTQ record");
                if (rec.header.e == rec.header.b || in.p[rec.header.b] != '@') fastqError(recordNumber + 1, "expected header line starting with '@'");
                if (rec.plus.e == rec.plus.b || in.p[rec.plus.b] != '+') fastqError(recordNumber + 1, "expected separator line starting with '+'");
                if (logicalLen(rec.sequence) != logicalLen(rec.quality)) fastqError(recordNumber + 1, "sequence and quality length differ");
                ++recordNumber;
                b->seq.insert(b->seq.end(), in.p + rec.sequence.b, in.p + rec.sequence.e);
                b->offsets.push_back(b->seq.size());
                b->recs.push_back(rec);
            }
            msPack += since(t0);
            if (!b->recs.empty()) {
                if (ctxThread.joinable()) {
                    ctxThread.join();
                    if (ctxRc != TLS_OK) die("Error: tls_create failed: %s\n", ctxErr.c_str());
                    timer.mark("first batch + CUDA context");
                }
                t0 = std::chrono::steady_clock::now();
                b->seq.resize(b->seq.size() + 256, 0);
                tls_batch tb{};
                tb.seq = b->seq.data(); tb.offsets = b->offsets.data(); tb.abs_pos = nullptr;
                tb.n_items = static_cast<uint32_t>(b->recs.size()); tb.mode = TLS_MODE_READS;
                check(tls_submit(ctx, &tb, &b->ticket), "tls_submit");
                msSubmit += since(t0);
                flight.push_back(std::move(b));
            }
            if (more && flight.size() < 2) continue;
        }
        if (!flight.empty()) {
            const auto t0 = std::chrono::steady_clock::now();
            consume(*flight.front());
            msCollect += since(t0);
            pool.push_back(std::move(flight.front()));
            flight.erase(flight.begin());
        }
    }
    timer.mark("scan");
    if (timer.on) fprintf(stderr, "[pipe]   of which: parse + pack %.1f ms, tls_submit %.1f ms, tls_collect + output %.1f ms\n", msPack, msSubmit, msCollect);
    if (!outBuf.empty()) fwrite(outBuf.data(), 1, outBuf.size(), out);
    fflush(out);
    if (out != stdout) {
        fclose(out);
        fprintf(stderr, "Wrote telomeric reads to %s.\n", outPath.c_str());
    }
    fprintf(stderr, "FASTQ subset: kept %llu of %llu reads.\n", static_cast<unsigned long long>(passedReads),
            static_cast<unsigned long long>(totalReads));
    fflush(stderr);
    _exit(EXIT_SUCCESS);
}

}  // namespace

int main(int argc, char **argv) {
    if (argc == 1) {
        printf("teloscope-b200 input.[fa|fa.gz] [options]\nUse -h for additional help.\n");
        return 0;
    }
    if (argc == 4 && !strcmp(argv[1], "--gfa-normalize")) return normalizeGfa(argv[2], argv[3]);   // developer / test entry: graph in, GFA 1.2 out, no scan
    Options o = parseArgs(argc, argv);
    if (o.inSequence.empty() && (o.bamSubset || o.fastqSubset) && !isatty(STDIN_FILENO)) {      // pipe-only invocation, src/main.cpp:567-573
        o.inSequenceName = "stdin";
        if (o.outRoute.empty()) o.outRoute = ".";
    } else if (o.inSequence.empty()) {
        if (!isatty(STDIN_FILENO)) unsupported("an assembly on stdin");
        die("Error: No input file provided. Use -f or pass as positional argument.\n");
    }
    if (o.fastqSubset && o.bamSubset) die("Error: --fastq-subset and --bam-subset are mutually exclusive.\n");
    if (o.step > o.windowSize) die("Error: Step size (%d) cannot be larger than window size (%d).\n", o.step, o.windowSize);
    if (o.step < o.windowSize) fprintf(stderr, "Sliding with window size (%d) and step size (%d).\n", o.windowSize, o.step);
    if (!o.fastqSubset && !o.bamSubset) {   // writable check, src/main.cpp:611-622
        const std::string testPath = o.outRoute + "/.teloscope_write_test";
        std::ofstream test(testPath);
        if (!test.is_open()) die("Error: Output directory '%s' is not writable.\n", o.outRoute.c_str());
        test.close();
        std::filesystem::remove(testPath);
    }
    if (!o.hasInputPatterns) o.rawPatterns = {o.canonicalFwd, o.canonicalRev};
    if (o.rawPatterns.empty()) {
        fprintf(stderr, "Warning: No valid patterns supplied via -p. Using canonical: %s, %s\n", o.canonicalFwd.c_str(), o.canonicalRev.c_str());
        o.rawPatterns = {o.canonicalFwd, o.canonicalRev};
    }

    // ---- pattern table (expandPatternsWithOrientation through the C ABI) ------------------------------------
    std::string csv;
    for (const auto &p : o.rawPatterns) { if (!csv.empty()) csv += ','; csv += p; }
    uint32_t nPat = 0;
    size_t textNeed = 0;
    check(tls_expand_patterns(csv.c_str(), o.editDistance, o.canonicalFwd.c_str(), nullptr, 0, nullptr, nullptr, 0, &nPat, &textNeed),
          "pattern expansion");
    std::vector<char> text(std::max<size_t>(textNeed, 1));
    std::vector<uint8_t> isFwd(std::max<uint32_t>(nPat, 1)), isCan(std::max<uint32_t>(nPat, 1));
    check(tls_expand_patterns(csv.c_str(), o.editDistance, o.canonicalFwd.c_str(), text.data(), text.size(), isFwd.data(), isCan.data(),
                              nPat, &nPat, &textNeed), "pattern expansion");
    std::vector<const char *> patPtr;
    for (size_t off = 0, i = 0; i < nPat; ++i) { patPtr.push_back(text.data() + off); off += strlen(text.data() + off) + 1; }
    fprintf(stderr, "Scanning %zu telomeric variants (includes reverse complements).\n", static_cast<size_t>(nPat));
    if (o.editDistance > 0)
        fprintf(stderr, "Edit distance enabled: up to %u substitution%s per seed.\n", o.editDistance, o.editDistance > 1 ? "s" : "");
    if (nPat > 500) {
        fprintf(stderr, "Warning: %zu patterns is unusually high and may be slow on large genomes.\n", static_cast<size_t>(nPat));
        fprintf(stderr, "  Consider fewer IUPAC wildcards or a lower -x value.\n");
    }
    {
        std::string summary;
        auto add = [&](bool on, const char *label) { if (on) { if (!summary.empty()) summary += ", "; summary += label; } };
        add(o.outGC, "GC windows"); add(o.outWinRepeats, "repeat density"); add(o.outEntropy, "Shannon entropy");
        add(o.outMatches, "genome-wide matches"); add(o.outITS, "ITS blocks");
        if (!summary.empty()) fprintf(stderr, "Outputs: %s.\n", summary.c_str());
    }
    if ((o.fastqSubset || o.bamSubset) && (o.outFasta || o.outWinRepeats || o.outGC || o.outEntropy || o.outMatches || o.outITS || o.manualCuration))
        fprintf(stderr, "Warning: assembly output flags are ignored in read subset mode.\n");
    if (o.cmdFlag) {
        FILE *cs = (o.fastqSubset || o.bamSubset) ? stderr : stdout;
        for (int i = 0; i < argc; ++i) fprintf(cs, "%s ", argv[i]);
        fprintf(cs, "\n");
    }

    StageTimer timer;
    const unsigned threads = o.threads ? o.threads : std::max(1u, std::thread::hardware_concurrency());

    // ---- device context: created on its own thread while the input is read and parsed ----------------------------------
    tls_config cfg{};
    cfg.abi_version = TLS_ABI_VERSION;
    cfg.n_patterns = nPat; cfg.patterns = patPtr.data(); cfg.is_forward = isFwd.data(); cfg.is_canonical = isCan.data();
    cfg.window_size = o.windowSize; cfg.step = o.step; cfg.terminal_limit = o.terminalLimit;
    cfg.max_match_dist = o.maxMatchDist; cfg.min_block_len = o.minBlockLen; cfg.max_block_dist = o.maxBlockDist;
    cfg.min_block_counts = o.minBlockCounts; cfg.min_block_density = o.minBlockDensity;
    cfg.outputs = ((o.outGC || o.outEntropy) ? 1u * TLS_OUT_ACGT : 0u) | (o.outMatches ? 1u * TLS_OUT_MATCHES : 0u) | TLS_OUT_ITS;
    if (o.fastqSubset) return runFastqSubset(o, cfg);
    if (o.bamSubset) return runBamSubset(o, cfg);
    if (isGfaPath(o.inSequence)) return runGfa(o, cfg);
    // One context per device (--gpus N: contexts 0 .. N-1 on devices TLS_DEVICE + k modulo the visible devices), each
    // used under its own lock: a feeder thread submits the device's batches, the main thread collects them.
    struct Dev { tls_ctx *ctx = nullptr; std::mutex mu; int inFlight = 0; };
    const int nDev = o.gpus;
    std::vector<Dev> devs(static_cast<size_t>(nDev));
    int ctxRc = TLS_OK;
    std::string ctxErr;
    std::thread ctxThread([&]() {
        const int visible = std::max(1, tls_device_count());
        const int first = getenv("TLS_DEVICE") ? atoi(getenv("TLS_DEVICE")) : 0;
        std::vector<std::thread> ts;
        std::vector<int> rcs(static_cast<size_t>(nDev), TLS_OK);
        std::vector<std::string> errs(static_cast<size_t>(nDev));
        for (int d = 0; d < nDev; ++d)
            ts.emplace_back([&, d]() {
                rcs[static_cast<size_t>(d)] = tls_create(&cfg, (first + d) % visible, &devs[static_cast<size_t>(d)].ctx);
                if (rcs[static_cast<size_t>(d)] != TLS_OK) errs[static_cast<size_t>(d)] = tls_last_error();     // the message is thread-local
            });
        for (auto &t : ts) t.join();
        for (int d = 0; d < nDev; ++d) if (rcs[static_cast<size_t>(d)] != TLS_OK && ctxRc == TLS_OK) { ctxRc = rcs[static_cast<size_t>(d)]; ctxErr = errs[static_cast<size_t>(d)]; }
    });

    // ---- ingest ---------------------------------------------------------------------------------------------
    std::vector<PathRec> paths;
    std::vector<Item> items;
    uint8_t *seqOwner = nullptr;
    uint64_t seqBytes = 0;
    {
        InputData in;
        loadInput(o.inSequence, in);
        timer.mark("read file");
        if (in.n > 1 && (in.p[0] == 'H' || in.p[0] == 'S') && in.p[1] == '\t') { ctxThread.join(); unsupported("GFA content in a file that is not named *.gfa / *.gfa.gz"); }
        if (in.n > 0 && in.p[0] == '@') { ctxThread.join(); unsupported("FASTQ input"); }
        if (in.n == 0 || in.p[0] != '>') { ctxThread.join(); die("Error: '%s' does not look like a FASTA file.\n", o.inSequence.c_str()); }
        parseFasta(in, threads, paths, items, seqOwner, seqBytes);
    }
    const uint8_t *seq = seqOwner;
    timer.mark("parse FASTA");
    ctxThread.join();
    if (ctxRc != TLS_OK) die("Error: tls_create failed: %s\n", ctxErr.c_str());
    timer.mark("tls_create (CUDA context)");

    // ---- output files (src/teloscope.cpp:886-920: which exist depends on the flags only) ---------------------------
    const std::string base = o.outRoute + "/" + o.inSequenceName;
    OutFile fDensity, fCanRatio, fStrand, fGC, fEntropy, fCan, fNon, fTerminal, fIts, fGaps, fReport;
    if (o.outWinRepeats) {
        fDensity.open(base + "_window_repeat_density.bedgraph");
        fCanRatio.open(base + "_window_canonical_ratio.bedgraph");
        fStrand.open(base + "_window_strand_ratio.bedgraph");
        fDensity.buf += "track type=bedGraph name=\"Repeat Density\" description=\"Total repeat density per window\"\n";
        fCanRatio.buf += "track type=bedGraph name=\"Canonical Ratio\" description=\"Canonical fraction of repeat density per window\"\n";
        fStrand.buf += "track type=bedGraph name=\"Strand Ratio\" description=\"Forward-strand fraction of repeat density per window\"\n";
    }
    if (o.outGC) { fGC.open(base + "_window_gc.bedgraph"); fGC.buf += "track type=bedGraph name=\"GC Content\" description=\"GC content per window\"\n"; }
    if (o.outEntropy) { fEntropy.open(base + "_window_entropy.bedgraph"); fEntropy.buf += "track type=bedGraph name=\"Shannon Entropy\" description=\"Shannon entropy per window\"\n"; }
    if (o.outMatches) { fCan.open(base + "_canonical_matches.bed"); fNon.open(base + "_noncanonical_matches.bed"); }
    if (o.outITS) fIts.open(base + "_interstitial_telomeres.bed");
    fTerminal.open(base + "_terminal_telomeres.bed");
    fGaps.open(base + "_gaps.bed");
    fReport.open(base + "_report.tsv");

    // ---- scan: batches of whole contigs, two in flight; results are consumed in submission order ---------------------
    const uint64_t batchBytes = getenv("TLS_BATCH_BYTES") ? strtoull(getenv("TLS_BATCH_BYTES"), nullptr, 10) : (256ull << 20);
    const uint32_t mode = o.ultraFastMode ? TLS_MODE_TIPS : TLS_MODE_GENOME;
    // The batch list is fixed up front: whole contigs, at most batchBytes each (smaller with several devices, so that every
    // device gets several), batch k on device k mod N.  Results are consumed in batch order = input order.
    struct Batch { size_t first = 0, count = 0; std::vector<uint64_t> offsets, absPos; int ticket = -1; int dev = 0; bool submitted = false; };
    double msSubmit = 0, msCollect = 0, msFormat = 0, msWrite = 0;
    auto since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    std::vector<Batch> batches;
    {
        uint64_t total = 0;
        for (const Item &it : items) total += it.len;
        const uint64_t cap = nDev > 1 ? std::max<uint64_t>(16ull << 20, std::min<uint64_t>(batchBytes, total / (4ull * static_cast<uint64_t>(nDev)) + 1)) : batchBytes;
        size_t nextItem = 0;
        while (nextItem < items.size()) {
            Batch b;
            b.first = nextItem;
            uint64_t bytes = 0;
            while (nextItem < items.size() && (bytes == 0 || bytes + items[nextItem].len <= cap) && nextItem - b.first < 0x7FFFFFFFu) {
                bytes += items[nextItem].len;
                ++nextItem;
            }
            b.count = nextItem - b.first;
            for (size_t i = 0; i < b.count; ++i) { b.offsets.push_back(items[b.first + i].offset); b.absPos.push_back(items[b.first + i].absPos); }
            b.offsets.push_back(items[b.first + b.count - 1].offset + items[b.first + b.count - 1].len);
            b.dev = static_cast<int>(batches.size() % static_cast<size_t>(nDev));
            batches.push_back(std::move(b));
        }
    }
    std::mutex flowMu;
    std::condition_variable flowCv;
    std::vector<std::thread> feeders;
    for (int d = 0; d < nDev; ++d)
        feeders.emplace_back([&, d]() {
            Dev &dev = devs[static_cast<size_t>(d)];
            for (size_t k = static_cast<size_t>(d); k < batches.size(); k += static_cast<size_t>(nDev)) {
                Batch &b = batches[k];
                { std::unique_lock<std::mutex> lk(flowMu); flowCv.wait(lk, [&] { return dev.inFlight < 2; }); ++dev.inFlight; }
                tls_batch tb{};
                tb.seq = seq; tb.offsets = b.offsets.data(); tb.abs_pos = b.absPos.data();
                tb.n_items = static_cast<uint32_t>(b.count); tb.mode = mode;
                const auto t0 = std::chrono::steady_clock::now();
                { std::lock_guard<std::mutex> lk(dev.mu); check(tls_submit(dev.ctx, &tb, &b.ticket), "tls_submit"); }
                { std::lock_guard<std::mutex> lk(flowMu); msSubmit += since(t0); b.submitted = true; }
                flowCv.notify_all();
            }
        });
    // Formatting is the host's share of the genome-wide outputs (3 M windows x 5 tracks on a human assembly): the
    // windows and match lists of a batch are cut into chunks, formatted in parallel into per-chunk buffers and
    // appended to the files in order.
    struct Chunk { size_t item; uint64_t lo, hi; int kind; std::string out[5]; };      // kind 0: windows, 1: canonical, 2: non-canonical
    auto consume = [&](const Batch &b) {
        Dev &dev = devs[static_cast<size_t>(b.dev)];
        tls_result r{};
        auto t0 = std::chrono::steady_clock::now();
        { std::lock_guard<std::mutex> lk(dev.mu); check(tls_collect(dev.ctx, b.ticket, &r), "tls_collect"); }
        msCollect += since(t0);
        t0 = std::chrono::steady_clock::now();
        std::vector<Chunk> chunks;
        constexpr uint64_t kChunk = 16384;
        for (size_t j = 0; j < b.count; ++j) {
            const Item &it = items[b.first + j];
            PathRec &p = paths[it.path];
            for (uint64_t k = r.terminal_off[j]; k < r.terminal_off[j + 1]; ++k) p.terminal.push_back(r.terminal[k]);
            if (mode != TLS_MODE_GENOME) continue;
            const size_t its0 = p.its.size();
            for (uint64_t k = r.its_off[j]; k < r.its_off[j + 1]; ++k) p.its.push_back(r.its[k]);
            std::sort(p.its.begin() + static_cast<long>(its0), p.its.end(), [](const tls_block &a, const tls_block &b2) { return a.start < b2.start; });
            p.nCanonical += r.n_canonical[j];
            p.nWindows += r.window_off[j + 1] - r.window_off[j];
            if (o.outWinRepeats || o.outEntropy || o.outGC)
                for (uint64_t k = r.window_off[j]; k < r.window_off[j + 1]; k += kChunk)
                    chunks.push_back({j, k, std::min(k + kChunk, r.window_off[j + 1]), 0, {}});
            if (o.outMatches) {
                for (uint64_t k = r.canonical_off[j]; k < r.canonical_off[j + 1]; k += kChunk)
                    chunks.push_back({j, k, std::min(k + kChunk, r.canonical_off[j + 1]), 1, {}});
                for (uint64_t k = r.noncanonical_off[j]; k < r.noncanonical_off[j + 1]; k += kChunk)
                    chunks.push_back({j, k, std::min(k + kChunk, r.noncanonical_off[j + 1]), 2, {}});
            }
        }
        parallelFor(chunks.size(), threads, [&](size_t ci) {
            Chunk &c = chunks[ci];
            const Item &it = items[b.first + c.item];
            const std::string &header = paths[it.path].header;
            if (c.kind == 0) {
                const uint64_t w0 = r.window_off[c.item];
                for (uint64_t k = c.lo; k < c.hi; ++k) {
                    const tls_window &w = r.windows[k];
                    const uint64_t rel = (k - w0) * o.step;
                    const uint32_t size = static_cast<uint32_t>(std::min<uint64_t>(o.windowSize, it.len - rel));
                    const uint64_t ws = it.absPos + rel, we = ws + size;
                    if (o.outWinRepeats) {                                         // src/teloscope.cpp:788-803
                        const uint32_t totalCovered = w.fwd_cov + w.rev_cov;
                        const float totalDensity = static_cast<float>(totalCovered) / size;
                        const float canonRatio = totalCovered > 0 ? static_cast<float>(w.can_cov) / (w.can_cov + w.noncan_cov) : -1.0f;
                        const float strandRatio = totalCovered > 0 ? static_cast<float>(w.fwd_cov) / (w.fwd_cov + w.rev_cov) : -1.0f;
                        windowLine(c.out[0], header, ws, we, totalDensity);
                        windowLine(c.out[1], header, ws, we, canonRatio);
                        windowLine(c.out[2], header, ws, we, strandRatio);
                    }
                    if (o.outEntropy) windowLine(c.out[3], header, ws, we, shannonEntropy(w.acgt, size));
                    if (o.outGC) windowLine(c.out[4], header, ws, we, gcContent(w.acgt, size));
                }
            } else {
                const tls_match *m = c.kind == 1 ? r.canonical : r.noncanonical;
                std::string &s = c.out[0];
                for (uint64_t k = c.lo; k < c.hi; ++k) {
                    s += header; s.push_back('\t'); appendU64(s, m[k].pos); s.push_back('\t'); appendU64(s, m[k].pos + m[k].len);
                    s.push_back('\t');
                    const uint8_t *src = seq + it.offset + (m[k].pos - it.absPos);
                    for (uint32_t q = 0; q < m[k].len; ++q) s.push_back(static_cast<char>(std::toupper(src[q])));
                    s.push_back('\n');
                }
            }
        });
        msFormat += since(t0);
        t0 = std::chrono::steady_clock::now();
        auto put = [](OutFile &f, const std::string &s) {
            if (s.empty() || !f.fh) return;
            f.flush();
            fwrite(s.data(), 1, s.size(), f.fh);
        };
        for (const Chunk &c : chunks) {
            if (c.kind == 0) { put(fDensity, c.out[0]); put(fCanRatio, c.out[1]); put(fStrand, c.out[2]); put(fEntropy, c.out[3]); put(fGC, c.out[4]); }
            else put(c.kind == 1 ? fCan : fNon, c.out[0]);
        }
        msWrite += since(t0);
        { std::lock_guard<std::mutex> lk(dev.mu); check(tls_release(dev.ctx, b.ticket), "tls_release"); }
        { std::lock_guard<std::mutex> lk(flowMu); --dev.inFlight; }
        flowCv.notify_all();
    };
    for (size_t k = 0; k < batches.size(); ++k) {
        { std::unique_lock<std::mutex> lk(flowMu); flowCv.wait(lk, [&] { return batches[k].submitted; }); }
        consume(batches[k]);
    }
    for (auto &t : feeders) t.join();
    timer.mark("scan + window / match files");
    if (timer.on) fprintf(stderr, "[pipe]   of which: tls_submit %.1f ms, tls_collect %.1f ms, formatting %.1f ms, file writes %.1f ms\n", msSubmit, msCollect, msFormat, msWrite);
    // the ctx is not destroyed: freeing gigabytes of device memory and unloading the runtime costs ~0.1 s that the
    // process exit (below: _exit after the files are closed) gives back to the driver anyway
    // ---- per path: classification, block / gap rows, report (src/teloscope.cpp:687-847) -------------------------------------
    std::string rep;           // goes to stdout and to the report file alike (after the banner)
    std::cout << "\n+++ Path Summary Report +++\n";
    rep += o.ultraFastMode ? "pos\theader\ttelomeres\tlabels\tgaps\ttype\tgranular\n"
                           : "pos\theader\ttelomeres\tlabels\tgaps\ttype\tgranular\tits\tcanonical\twindows\n";
    uint32_t totalPaths = static_cast<uint32_t>(paths.size()), totalNWindows = 0, totalTelomeres = 0, totalITS = 0, totalCanMatches = 0, totalGaps = 0;
    uint32_t typeCount[10] = {0};
    std::vector<float> telomereLengths;
    std::vector<uint64_t> scaffoldLens, contigLens;
    for (PathRec &p : paths) {
        labelTerminalBlocks(p, o.terminalLimit);
        int longestCount = 0;
        std::string longestLabels;
        for (size_t i = 0; i < p.terminal.size(); ++i) {
            const tls_block &b = p.terminal[i];
            const uint64_t blockEnd = b.start + b.len;
            const bool scaffoldTerminal = b.start < o.terminalLimit || blockEnd > p.pathSize - o.terminalLimit;
            if (scaffoldTerminal || o.manualCuration) {
                blockRow(fTerminal.buf, p.header, b, p.pathSize);
                fTerminal.buf += scaffoldTerminal ? "\tscaffold\n" : "\tcontig\n";
            }
            if (p.isLongest[i]) {
                ++longestCount;
                longestLabels += b.label;
                telomereLengths.push_back(static_cast<float>(b.len));
            }
        }
        if (o.outITS)
            for (const tls_block &b : p.its) { blockRow(fIts.buf, p.header, b, p.pathSize); fIts.buf.push_back('\n'); }
        for (const Gap &g : p.gaps) {
            fGaps.buf += p.header; fGaps.buf.push_back('\t'); appendU64(fGaps.buf, g.start); fGaps.buf.push_back('\t');
            appendU64(fGaps.buf, g.start + g.length); fGaps.buf.push_back('\n');
        }
        fTerminal.flushIfBig(); fIts.flushIfBig(); fGaps.flushIfBig();
        const uint16_t gaps = static_cast<uint16_t>(p.gaps.size());
        appendU64(rep, static_cast<uint64_t>(p.seqPos) + 1); rep.push_back('\t'); rep += p.header; rep.push_back('\t');
        appendU64(rep, static_cast<uint64_t>(longestCount)); rep.push_back('\t'); rep += longestLabels.empty() ? "none" : longestLabels;
        rep.push_back('\t'); appendU64(rep, gaps); rep.push_back('\t'); rep += kTypeNames[p.scaffoldType]; rep.push_back('\t');
        rep += p.terminalLabel;
        totalTelomeres += static_cast<uint32_t>(longestCount);
        totalGaps += gaps;
        if (!o.ultraFastMode) {
            rep.push_back('\t'); appendU64(rep, p.its.size()); rep.push_back('\t'); appendU64(rep, p.nCanonical); rep.push_back('\t');
            appendU64(rep, p.nWindows);
            totalNWindows += static_cast<uint32_t>(p.nWindows);
            totalITS += static_cast<uint32_t>(p.its.size());
            totalCanMatches += static_cast<uint32_t>(p.nCanonical);
        }
        rep.push_back('\n');
        ++typeCount[p.scaffoldType];
        scaffoldLens.push_back(p.pathSize);                                      // computeSummaryCounts, :950-984
        std::vector<Gap> gs = p.gaps;
        std::sort(gs.begin(), gs.end(), [](const Gap &a, const Gap &b) { return a.start < b.start; });
        uint64_t prevEnd = 0;
        for (const Gap &g : gs) {
            if (g.start > prevEnd) contigLens.push_back(g.start - prevEnd);
            prevEnd = g.start + g.length;
        }
        if (p.pathSize > prevEnd) contigLens.push_back(p.pathSize - prevEnd);
    }
    float teloMean = 0.0f, teloMedian = 0.0f, teloMin = 0.0f, teloMax = 0.0f;
    if (!telomereLengths.empty()) {                                               // getStats, src/tools.cpp:22-51
        std::vector<float> &v = telomereLengths;
        float sum = 0.0f;
        teloMin = teloMax = v[0];
        for (float x : v) { if (x < teloMin) teloMin = x; if (x > teloMax) teloMax = x; sum += x; }
        teloMean = sum / v.size();
        const size_t mid = v.size() / 2;
        std::nth_element(v.begin(), v.begin() + static_cast<long>(mid), v.end());
        teloMedian = v.size() % 2 == 0 ? (v[mid] + *std::max_element(v.begin(), v.begin() + static_cast<long>(mid))) / 2 : v[mid];
    }
    const uint64_t scaffoldN50 = computeN50(scaffoldLens), contigN50 = computeN50(contigLens);
    auto kv = [&](const char *k, uint64_t v) { rep += k; appendU64(rep, v); rep.push_back('\n'); };
    auto kf = [&](const char *k, float v) { rep += k; appendFloat(rep, v); rep.push_back('\n'); };
    rep += "\n+++ Assembly Summary Report +++\n";
    kv("Total paths:\t", totalPaths); kv("Total gaps:\t", totalGaps); kv("Scaffold N50:\t", scaffoldN50); kv("Contig N50:\t", contigN50);
    kv("Total telomeres:\t", totalTelomeres);
    if (!o.ultraFastMode) { kv("Total ITS blocks:\t", totalITS); kv("Total canonical matches:\t", totalCanMatches); kv("Total windows analyzed:\t", totalNWindows); }
    rep += "\n+++ Telomere Statistics +++\n";
    if (totalTelomeres > 0) { kf("Mean length:\t", teloMean); kf("Median length:\t", teloMedian); kf("Min length:\t", teloMin); kf("Max length:\t", teloMax); }
    else rep += "No telomeres found for statistics.\n";
    rep += "\n+++ Chromosome Telomere Counts+++\n";
    kv("Two telomeres:\t", typeCount[0] + typeCount[1] + typeCount[2] + typeCount[3]);
    kv("One telomere:\t", typeCount[4] + typeCount[5]);
    kv("Zero telomeres:\t", typeCount[6] + typeCount[7]);
    rep += "\n+++ Chromosome Telomere/Gap Completeness+++\n";
    kv("T2T:\t", typeCount[0]); kv("Gapped T2T:\t", typeCount[1]); kv("Misassembled:\t", typeCount[2]); kv("Gapped misassembled:\t", typeCount[3]);
    kv("Incomplete:\t", typeCount[4]); kv("Gapped incomplete:\t", typeCount[5]); kv("No telomeres:\t", typeCount[6]);
    kv("Gapped no telomeres:\t", typeCount[7]); kv("Discordant:\t", typeCount[8]); kv("Gapped discordant:\t", typeCount[9]);
    std::cout << rep;
    std::cout.flush();
    fReport.buf += rep;
    for (OutFile *f : {&fDensity, &fCanRatio, &fStrand, &fGC, &fEntropy, &fCan, &fNon, &fTerminal, &fIts, &fGaps, &fReport}) f->close();
    timer.mark("blocks, report, close");
    fflush(stdout);
    fflush(stderr);
    _exit(EXIT_SUCCESS);
}
