/*
 * oracle/oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A from-scratch CPU restatement of the Teloscope scan core, used only as the checker for the
 * CUDA path (tests/, __graft_entry__.smoke(), bench.py cpu_baseline).  Every function cites the
 * reference file:line whose behaviour it restates (paths relative to /root/reference).
 * Parity status: pinned against the compiled reference (see oracle.h).
 */
#include "oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Pattern expansion: src/tools.cpp:54-127 (IUPAC + substitution variants), :201-283 (orientation)
// ---------------------------------------------------------------------------------------------

const char *iupacSet(char c) {  // src/tools.cpp:54-70
    switch (c) {
        case 'A': return "A";   case 'C': return "C";   case 'G': return "G";   case 'T': return "T";
        case 'R': return "AG";  case 'Y': return "CT";  case 'M': return "AC";  case 'K': return "GT";
        case 'S': return "CG";  case 'W': return "AT";  case 'H': return "ACT"; case 'B': return "CGT";
        case 'V': return "ACG"; case 'D': return "AGT"; case 'N': return "ACGT";
        default:  return "";    // unknown letter: operator[] yields an empty vector -> no combos
    }
}

void iupacCombos(const std::string &seed, std::string &cur, size_t i, std::vector<std::string> &out) {
    if (i == seed.size()) { out.push_back(cur); return; }      // src/tools.cpp:73-84
    for (const char *p = iupacSet(seed[i]); *p; ++p) {
        cur[i] = *p;
        iupacCombos(seed, cur, i + 1, out);
    }
}

bool isACGT(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

std::vector<std::string> substitutionVariants(const std::string &s, int maxDist) {
    std::vector<std::string> v;                                 // src/tools.cpp:87-127
    if (maxDist <= 0) return v;
    static const char nts[4] = {'A', 'C', 'G', 'T'};
    for (size_t i = 0; i < s.size(); ++i) {
        if (!isACGT(s[i])) continue;
        for (char c : nts) if (c != s[i]) { std::string t = s; t[i] = c; v.push_back(t); }
    }
    if (maxDist >= 2) {                                         // distance 2 = variants of variants
        size_t n1 = v.size();
        for (size_t k = 0; k < n1; ++k) {
            std::vector<std::string> d2 = substitutionVariants(v[k], 1);
            v.insert(v.end(), d2.begin(), d2.end());
        }
    }
    return v;
}

bool revComp(const std::string &s, std::string &out) {         // gfalibs/include/functions.h:348-382
    out.assign(s.rbegin(), s.rend());
    for (char &c : out) {
        switch (c) {
            case 'A': c = 'T'; break; case 'C': c = 'G'; break;
            case 'G': c = 'C'; break; case 'T': c = 'A'; break;
            case 'a': c = 't'; break; case 'c': c = 'g'; break;
            case 'g': c = 'c'; break; case 't': c = 'a'; break;
            case 'N': case 'n': case 'X': case 'x': case '*': break;
            default: return false;                              // reference throws domain_error
        }
    }
    return true;
}

unsigned bestUngappedDist(const std::string &shorter, const std::string &longer) {
    unsigned best = 255;                                        // src/tools.cpp:222-244
    for (size_t off = 0; off + shorter.size() <= longer.size(); ++off) {
        unsigned d = 0;
        for (size_t i = 0; i < shorter.size(); ++i) d += shorter[i] != longer[off + i];
        best = std::min<unsigned>(best, d & 0xFF);
    }
    return best;
}

// returns false on invalid nucleotide
bool expandWithOrientation(const std::vector<std::string> &raw, int x, const std::string &canFwd,
                           std::vector<std::pair<std::string, bool>> &out) {
    std::string canRev;
    if (!revComp(canFwd, canRev)) return false;
    for (const std::string &seed : raw) {
        if (seed.empty()) continue;
        std::vector<std::string> combos;
        std::string work = seed;
        iupacCombos(seed, work, 0, combos);
        for (const std::string &combo : combos) {
            bool fwd;                                           // isCloserToFwd, src/tools.cpp:206-245
            if (combo.size() == canFwd.size()) {
                unsigned df = 0, dr = 0;
                for (size_t i = 0; i < combo.size(); ++i) { df += combo[i] != canFwd[i]; dr += combo[i] != canRev[i]; }
                fwd = (df & 0xFF) <= (dr & 0xFF);
            } else {
                bool patShorter = combo.size() < canFwd.size();
                std::string comboRc;
                if (!revComp(combo, comboRc)) return false;
                const std::string &shorter = patShorter ? combo : canFwd;
                const std::string &longer = patShorter ? canFwd : combo;
                const std::string &longerRev = patShorter ? canRev : comboRc;
                fwd = bestUngappedDist(shorter, longer) <= bestUngappedDist(shorter, longerRev);
            }
            std::vector<std::string> variants{combo};
            std::vector<std::string> edits = substitutionVariants(combo, x);
            variants.insert(variants.end(), edits.begin(), edits.end());
            for (const std::string &v : variants) {
                std::string rc;
                if (!revComp(v, rc)) return false;
                out.emplace_back(v, fwd);
                out.emplace_back(rc, !fwd);
            }
        }
    }
    // src/tools.cpp:275-280: std::sort by string (unstable) then unique by string, first kept.
    // The survivor's flag is order dependent only if one string occurs with both flags; we use
    // the same std::sort on the same element sequence so libstdc++ behaves identically.
    std::sort(out.begin(), out.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    out.erase(std::unique(out.begin(), out.end(),
                          [](const auto &a, const auto &b) { return a.first == b.first; }),
              out.end());
    return true;
}

// ---------------------------------------------------------------------------------------------
// Trie: include/teloscope.h:15-86
// ---------------------------------------------------------------------------------------------

struct Node { int32_t kid[4] = {-1, -1, -1, -1}; bool end = false, fwd = false, can = false; };

int baseIdx(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; } }

struct Matcher {
    std::vector<Node> nodes{Node{}};
    unsigned longest = 0;
    void add(const std::string &p, bool fwd, bool can) {
        int32_t cur = 0;
        for (char ch : p) {
            int b = baseIdx(ch);
            if (b < 0) continue;                                // include/teloscope.h:44
            if (nodes[cur].kid[b] < 0) { nodes[cur].kid[b] = (int32_t)nodes.size(); nodes.emplace_back(); }
            cur = nodes[cur].kid[b];
        }
        nodes[cur].end = true; nodes[cur].fwd = fwd; nodes[cur].can = can;
        longest = std::max<unsigned>(longest, (unsigned)p.size());
    }
    int32_t step(int32_t n, char ch) const { int b = baseIdx(ch); return b < 0 ? -1 : nodes[n].kid[b]; }
};

struct Match { uint64_t pos; uint16_t len; bool fwd, can; };
struct Block {
    uint64_t start = 0; uint32_t len = 0, counts = 0, fwdCount = 0, revCount = 0, canCount = 0,
    nonCanCount = 0, totalCov = 0, fwdCov = 0, canCov = 0; bool validOr = true; char label = 0;
};
struct Window {
    uint64_t start = 0; uint32_t size = 0; uint32_t nt[4] = {0, 0, 0, 0};
    uint32_t canCov = 0, nonCanCov = 0, fwdCov = 0, revCov = 0; float gc = 0.f, entropy = 0.f;
};

struct Segment {
    std::vector<Window> windows;
    std::vector<Block> terminal, its;
    std::vector<Match> canonical, nonCanonical, fwd, rev, all;
};

}  // namespace

struct orc_ctx {
    Matcher trie;
    orc_params p;
    uint16_t itsMinLen = 0;     // 2 * patterns.front().size(), src/teloscope.cpp:185
};

struct orc_result {
    std::vector<orc_match> m[5];
    std::vector<orc_block> b[2];
    std::vector<orc_window> w;
};

namespace {

// Running accumulator shared by the terminal and interstitial run builders
// (src/teloscope.cpp:57-73 and :193-209 hold the same nine variables).
struct Run {
    bool open = false;
    uint64_t start = 0, end = 0, prev = 0;
    uint32_t n = 0, nFwd = 0, nCan = 0, cov = 0, fwdCov = 0, canCov = 0;
    void begin(const Match &m) {
        start = m.pos; end = m.pos + m.len; prev = m.pos; n = 1; nFwd = m.fwd; nCan = m.can;
        cov = m.len; fwdCov = m.fwd * (uint32_t)m.len; canCov = m.can * (uint32_t)m.len; open = true;
    }
    void absorb(const Match &m) {
        ++n; nFwd += m.fwd; nCan += m.can; cov += m.len;
        fwdCov += m.fwd * (uint32_t)m.len; canCov += m.can * (uint32_t)m.len; prev = m.pos;
    }
    Block toBlock() const {
        Block b; b.start = start; b.len = (uint32_t)(end - start); b.counts = n; b.fwdCount = nFwd;
        b.revCount = n - nFwd; b.canCount = nCan; b.nonCanCount = n - nCan; b.totalCov = cov;
        b.fwdCov = fwdCov; b.canCov = canCov; return b;
    }
};

// Teloscope::getTerminalBlocks, src/teloscope.cpp:29-176
uint64_t terminalBlocks(const orc_params &p, const std::vector<Match> &ms, std::vector<Block> &out,
                        uint64_t segSize, uint64_t absPos, bool fromStart) {
    uint64_t boundary = fromStart ? absPos : absPos + segSize;
    const int64_t n = (int64_t)ms.size();
    const uint32_t T = p.terminal_limit;
    auto inZone = [&](uint64_t pos) {                            // :48-52
        uint64_t rel = pos - absPos;
        if (segSize <= T) return true;
        return fromStart ? rel < T : rel >= segSize - T;
    };
    std::vector<Block> subs;
    Run r;
    auto closeRun = [&]() {                                      // :75-92
        // float32 compare: uint32 -> float on the left, float * (uint64 -> float) on the right
        float rhs = p.min_block_density * (float)(r.end - r.start);
        if (r.n >= p.min_block_counts && r.nCan > 0 && (float)r.canCov >= rhs) subs.push_back(r.toBlock());
        r.open = false;
    };
    const int64_t step = fromStart ? 1 : -1;
    for (int64_t i = fromStart ? 0 : n - 1; i != (fromStart ? n : -1); i += step) {   // :94-124
        const Match &m = ms[(size_t)i];
        if (!r.open) {
            if (!inZone(m.pos)) break;
            r.begin(m);
            continue;
        }
        uint64_t gap = fromStart ? m.pos - r.prev : r.prev - m.pos;
        if (gap <= p.max_match_dist) {
            if (fromStart) r.end = m.pos + m.len; else r.start = m.pos;
            r.absorb(m);
        } else {
            closeRun();
            if (!inZone(m.pos)) break;
            r.begin(m);
        }
    }
    if (r.open) closeRun();
    if (subs.empty()) return boundary;

    Block cur = subs[0];                                         // Phase 2, :130-174
    auto emit = [&]() {
        if (cur.len < p.min_block_len) return;
        cur.label = fromStart ? 'p' : 'q';
        uint64_t relStart = cur.start - absPos, relEnd = relStart + cur.len;
        uint64_t left = relStart, right = relEnd <= segSize ? segSize - relEnd : 0;
        cur.validOr = fromStart ? left <= right : left >= right;
        out.push_back(cur);
        boundary = fromStart ? cur.start + cur.len : cur.start;
    };
    for (size_t i = 1; i < subs.size(); ++i) {
        const Block &nx = subs[i];
        uint64_t gap = fromStart ? nx.start - (cur.start + cur.len) : cur.start - (nx.start + nx.len);
        if (gap <= p.max_block_dist) {
            if (fromStart) cur.len = (uint32_t)((nx.start + nx.len) - cur.start);
            else { cur.len = (uint32_t)((cur.start + cur.len) - nx.start); cur.start = nx.start; }
            cur.counts += nx.counts; cur.fwdCount += nx.fwdCount; cur.revCount += nx.revCount;
            cur.canCount += nx.canCount; cur.nonCanCount += nx.nonCanCount; cur.totalCov += nx.totalCov;
            cur.fwdCov += nx.fwdCov; cur.canCov += nx.canCov;
        } else {
            emit();
            cur = nx;
        }
    }
    emit();
    return boundary;
}

char blockLabel(uint32_t fwdCount, uint32_t counts) {           // include/teloscope.h:217-222
    float ratio = (fwdCount * 100.0f) / counts;
    if (ratio > 66.6f) return 'p';
    if (ratio < 33.3f) return 'q';
    return 'b';
}

// Teloscope::getInterstitialBlocks, src/teloscope.cpp:179-256
void interstitialBlocks(const orc_ctx &c, const std::vector<Match> &all, std::vector<Block> &out,
                        uint64_t fwdB, uint64_t revB) {
    auto it = std::lower_bound(all.begin(), all.end(), fwdB,
                               [](const Match &m, uint64_t v) { return m.pos < v; });
    if (it == all.end() || it->pos >= revB) return;
    Run r;
    auto closeRun = [&]() {
        Block b = r.toBlock();
        b.label = blockLabel(r.nFwd, r.n);
        if (b.len >= c.itsMinLen && r.nCan >= 4 && !(b.label == 'b' && r.nFwd < 2 && (r.n - r.nFwd) < 2))
            out.push_back(b);
        r.open = false;
    };
    for (; it != all.end() && it->pos < revB; ++it) {
        if (!r.open) r.begin(*it);
        else if (it->pos - r.prev <= c.p.max_match_dist) { r.end = it->pos + it->len; r.absorb(*it); }
        else { closeRun(); r.begin(*it); }
    }
    if (r.open) closeRun();
}

float gcContent(const uint32_t nt[4], uint32_t size) {          // include/teloscope.h:211-214
    uint32_t gc = nt[1] + nt[2];
    return static_cast<float>(gc) / size * 100.0;
}

float shannon(const uint32_t nt[4], uint32_t size) {            // include/teloscope.h:199-208
    float e = 0.0;
    for (int i = 0; i < 4; ++i)
        if (nt[i] > 0) { float pr = static_cast<float>(nt[i]) / size; e -= pr * std::log2(pr); }
    return std::round(e * 1000.0f) / 1000.0f;
}

// Teloscope::analyzeWindow, src/teloscope.cpp:387-534.  `win` is the window text, `carry` the
// counters that the next window inherits when step < windowSize.
void analyzeWindow(const orc_ctx &c, const char *win, uint32_t winLen, uint64_t winStart, Window &wd,
                   Window &carry, Segment &sd, uint64_t segSize, uint64_t absPos) {
    const uint32_t L = c.trie.longest;
    const uint32_t step = c.p.step, overlap = c.p.window_size - step, T = c.p.terminal_limit;
    const uint64_t winEnd = winStart + winLen;
    const uint64_t termEnd = segSize > T ? segSize - T : 0;      // :404
    const bool fullyTerm = winEnd <= T || winStart >= termEnd;
    const bool fullyInter = winStart > T && winEnd < termEnd;
    const bool mainAlways = overlap == 0 || winStart == 0;       // :409
    const bool hasOverlap = overlap != 0;
    // :413-415, evaluated in uint32 like the reference (may wrap when step or overlap < L)
    const uint32_t first = mainAlways ? 0u : std::min<uint32_t>(step - L, overlap - L);

    for (uint32_t i = first; i < winLen; ++i) {
        if (c.p.count_acgt) {                                    // :418-435
            int b = baseIdx(win[i]);
            if (b < 0) continue;
            if (mainAlways || i >= overlap) wd.nt[b]++;
            if (hasOverlap && i >= step) carry.nt[b]++;
        }
        int32_t node = 0;
        uint32_t limit = std::min<uint32_t>(i + L, winLen);      // :439 matches never cross window end
        for (uint32_t j = i; j < limit; ++j) {
            node = c.trie.step(node, win[j]);
            if (node < 0) break;
            const Node &nd = c.trie.nodes[(size_t)node];
            if (!nd.end) continue;
            Match m{absPos + winStart + i, (uint16_t)(j - i + 1), nd.fwd, nd.can};
            bool terminal;
            if (fullyTerm) terminal = true;
            else if (fullyInter) terminal = false;
            else { uint64_t a = winStart + i; terminal = a <= T || a >= termEnd; }   // :457-458 (<=)
            if (mainAlways || j >= overlap) {                    // :485-509
                if (m.can) { wd.canCov += m.len; sd.canonical.push_back(m); }
                else { wd.nonCanCov += m.len; if (terminal) sd.nonCanonical.push_back(m); }
                if (m.fwd) { wd.fwdCov += m.len; sd.fwd.push_back(m); }
                else { wd.revCov += m.len; sd.rev.push_back(m); }
                sd.all.push_back(m);
            }
            if (hasOverlap && i >= step) {                       // :512-529
                if (m.can) carry.canCov += m.len; else carry.nonCanCov += m.len;
                if (m.fwd) carry.fwdCov += m.len; else carry.revCov += m.len;
            }
        }
    }
}

// Teloscope::scanSegment, src/teloscope.cpp:537-658
Segment scanSegment(const orc_ctx &c, const char *seq, uint64_t n, uint64_t absPos, bool tipsOnly) {
    Segment sd;
    const uint32_t T = c.p.terminal_limit;
    const uint32_t L = c.trie.longest;
    if (tipsOnly) {
        auto region = [&](uint64_t a, uint64_t b) {              // :546-574
            for (uint64_t i = a; i < b; ++i) {
                int32_t node = 0;
                uint64_t limit = std::min<uint64_t>(i + L, b);   // truncated at the region end
                for (uint64_t j = i; j < limit; ++j) {
                    node = c.trie.step(node, seq[j]);
                    if (node < 0) break;
                    const Node &nd = c.trie.nodes[(size_t)node];
                    if (!nd.end) continue;
                    Match m{absPos + i, (uint16_t)(j - i + 1), nd.fwd, nd.can};
                    (nd.fwd ? sd.fwd : sd.rev).push_back(m);
                }
            }
        };
        if (n > (uint32_t)(2 * T)) { region(0, T); region(n - T, n); }   // :576 (2*T in uint32)
        else region(0, n);
    } else {
        const uint32_t W = c.p.window_size, S = c.p.step;        // :587-639
        Window prevCarry, nextCarry;
        uint64_t ws = 0;
        uint64_t cur = std::min<uint64_t>(W, n);
        while (ws < n) {
            Window wd = prevCarry;
            analyzeWindow(c, seq + ws, (uint32_t)cur, ws, wd, nextCarry, sd, n, absPos);
            wd.gc = gcContent(wd.nt, (uint32_t)cur);
            wd.entropy = shannon(wd.nt, (uint32_t)cur);
            wd.start = ws + absPos;
            wd.size = (uint32_t)cur;
            sd.windows.push_back(wd);
            prevCarry = nextCarry;
            nextCarry = Window();
            ws += S;
            if (ws >= n) break;
            cur = std::min<uint64_t>(W, n - ws);
        }
    }
    uint64_t fwdB = absPos, revB = absPos + n;                   // :643-655
    if (sd.fwd.size() >= 2) fwdB = terminalBlocks(c.p, sd.fwd, sd.terminal, n, absPos, true);
    if (sd.rev.size() >= 2) revB = terminalBlocks(c.p, sd.rev, sd.terminal, n, absPos, false);
    if (!tipsOnly && fwdB < revB && sd.all.size() >= 2) interstitialBlocks(c, sd.all, sd.its, fwdB, revB);
    return sd;
}

orc_match toC(const Match &m) { return orc_match{m.pos, m.len, (uint8_t)m.fwd, (uint8_t)m.can, 0}; }
orc_block toC(const Block &b) {
    return orc_block{b.start, b.len, b.counts, b.fwdCount, b.revCount, b.canCount, b.nonCanCount,
                     b.totalCov, b.fwdCov, b.canCov, (uint8_t)b.validOr, b.label, 0};
}

}  // namespace

extern "C" {

int orc_expand_patterns(const char *raw_csv, int x, const char *canFwd, char *out, size_t outLen) {
    std::vector<std::string> raw;
    std::string cur;
    for (const char *p = raw_csv;; ++p) {
        if (*p == ',' || *p == 0) { if (!cur.empty()) raw.push_back(cur); cur.clear(); if (!*p) break; }
        else cur.push_back(*p);
    }
    std::vector<std::pair<std::string, bool>> pats;
    if (!expandWithOrientation(raw, x, canFwd, pats)) return -2;
    std::string s;
    for (auto &pr : pats) { s += pr.first; s += '\t'; s += pr.second ? '1' : '0'; s += '\n'; }
    if (s.size() + 1 > outLen) return -1;
    std::memcpy(out, s.c_str(), s.size() + 1);
    return (int)pats.size();
}

orc_ctx *orc_create(const char *patterns, const uint8_t *fwd, int n, const char *canFwd,
                    const char *canRev, const orc_params *params) {
    orc_ctx *c = new orc_ctx;
    c->p = *params;
    const char *p = patterns;
    for (int i = 0; i < n; ++i) {
        std::string s(p);
        p += s.size() + 1;
        if (i == 0) c->itsMinLen = (uint16_t)(2 * s.size());
        c->trie.add(s, fwd[i] != 0, s == canFwd || s == canRev);   // include/teloscope.h:242-246
    }
    return c;
}

void orc_destroy(orc_ctx *c) { delete c; }

orc_result *orc_scan_segment(const orc_ctx *c, const char *seq, uint64_t n, uint64_t absPos, int tipsOnly) {
    Segment sd = scanSegment(*c, seq, n, absPos, tipsOnly != 0);
    orc_result *r = new orc_result;
    const std::vector<Match> *src[5] = {&sd.canonical, &sd.nonCanonical, &sd.fwd, &sd.rev, &sd.all};
    for (int k = 0; k < 5; ++k) for (const Match &m : *src[k]) r->m[k].push_back(toC(m));
    for (const Block &b : sd.terminal) r->b[0].push_back(toC(b));
    for (const Block &b : sd.its) r->b[1].push_back(toC(b));
    for (const Window &w : sd.windows) {
        orc_window o{};
        o.start = w.start; o.size = w.size;
        for (int i = 0; i < 4; ++i) o.acgt[i] = w.nt[i];
        o.can_cov = w.canCov; o.noncan_cov = w.nonCanCov; o.fwd_cov = w.fwdCov; o.rev_cov = w.revCov;
        o.gc = w.gc; o.entropy = w.entropy;
        r->w.push_back(o);
    }
    return r;
}

void orc_result_free(orc_result *r) { delete r; }

size_t orc_result_matches(const orc_result *r, int which, const orc_match **out) {
    *out = r->m[which].data();
    return r->m[which].size();
}
size_t orc_result_blocks(const orc_result *r, int which, const orc_block **out) {
    *out = r->b[which].data();
    return r->b[which].size();
}
size_t orc_result_windows(const orc_result *r, const orc_window **out) {
    *out = r->w.data();
    return r->w.size();
}

int orc_read_matches(const orc_ctx *c, const char *seq, uint64_t n) {
    orc_ctx rc = *c;                                             // src/read-filter.cpp:12-31
    rc.p.terminal_limit = std::numeric_limits<uint32_t>::max() / 2;
    std::string s(seq, n);
    if (!s.empty() && s.back() == '\r') s.pop_back();            // :38-40
    for (char &ch : s) if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
    Segment sd = scanSegment(rc, s.data(), s.size(), 0, true);
    return !sd.terminal.empty();
}

void orc_window_floats(const uint32_t acgt[4], uint32_t size, float *gc, float *entropy) {
    *gc = gcContent(acgt, size);
    *entropy = shannon(acgt, size);
}

uint64_t orc_scan_many(const orc_ctx *c, const char *buf, const uint64_t *off, int count,
                       int tipsOnly, int threads) {
    std::atomic<int> next{0};
    std::atomic<uint64_t> sum{0};
    auto work = [&]() {
        uint64_t local = 0;
        for (int i; (i = next.fetch_add(1)) < count;) {
            Segment sd = scanSegment(*c, buf + off[i], off[i + 1] - off[i], 0, tipsOnly != 0);
            local += sd.windows.size() + 3 * sd.all.size() + 5 * sd.fwd.size() + 7 * sd.rev.size() +
                     11 * sd.terminal.size() + 13 * sd.its.size() + 17 * sd.canonical.size();
            for (const Window &w : sd.windows) local += w.nt[0] + 2 * w.nt[1] + 3 * w.fwdCov;
        }
        sum += local;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    return sum.load();
}

}  // extern "C"
