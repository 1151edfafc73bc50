#include "patterns.hpp"

#include <algorithm>
#include <utility>

namespace tls {
namespace {

// IUPAC code -> bases (reference table: src/tools.cpp:54-70).  Unknown letters expand to nothing.
const char *iupac(char c) {
    static const char *const tbl[26] = {
        /*A*/ "A", /*B*/ "CGT", /*C*/ "C", /*D*/ "AGT", nullptr, nullptr, /*G*/ "G", /*H*/ "ACT",
        nullptr, nullptr, /*K*/ "GT", nullptr, /*M*/ "AC", /*N*/ "ACGT", nullptr, nullptr, nullptr,
        /*R*/ "AG", /*S*/ "CG", /*T*/ "T", nullptr, /*V*/ "ACG", /*W*/ "AT", nullptr, /*Y*/ "CT", nullptr};
    if (c < 'A' || c > 'Z') return "";
    const char *s = tbl[c - 'A'];
    return s ? s : "";
}

// Cartesian product of the per-position base sets, first position slowest — the order the
// reference's recursive getCombinations produces (src/tools.cpp:73-84).
std::vector<std::string> iupacExpand(const std::string &seed) {
    std::vector<std::string> acc{std::string()};
    for (char c : seed) {
        const char *set = iupac(c);
        std::vector<std::string> next;
        next.reserve(acc.size() * 4);
        for (const std::string &prefix : acc)
            for (const char *b = set; *b; ++b) next.push_back(prefix + *b);
        acc.swap(next);
    }
    return acc;
}

bool reverseComplement(const std::string &s, std::string &out) {
    out.resize(s.size());
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[s.size() - 1 - i], r;
        switch (c) {
            case 'A': r = 'T'; break;
            case 'C': r = 'G'; break;
            case 'G': r = 'C'; break;
            case 'T': r = 'A'; break;
            case 'N': case 'X': case '*': r = c; break;
            default: return false;
        }
        out[i] = r;
    }
    return true;
}

// All strings at substitution distance 1 (positions left to right, replacement bases in ACGT
// order), then — for distance 2 — the distance-1 neighbours of each of those, appended in order
// (src/tools.cpp:87-127).  Duplicates are kept; the caller de-duplicates after sorting.
void appendNeighbours(const std::string &s, std::vector<std::string> &out) {
    for (size_t i = 0; i < s.size(); ++i) {
        const char orig = s[i];
        if (orig != 'A' && orig != 'C' && orig != 'G' && orig != 'T') continue;
        for (char b : {'A', 'C', 'G', 'T'}) {
            if (b == orig) continue;
            out.push_back(s);
            out.back()[i] = b;
        }
    }
}

unsigned hammingAtBestOffset(const std::string &shorter, const std::string &longer) {
    unsigned best = 255;
    for (size_t off = 0; off + shorter.size() <= longer.size(); ++off) {
        unsigned d = 0;
        for (size_t i = 0; i < shorter.size(); ++i) d += shorter[i] != longer[off + i];
        best = std::min(best, d & 0xFFu);       // the reference counts in uint8_t
    }
    return best;
}

}  // namespace

bool expandPatterns(const std::vector<std::string> &rawPatterns, int editDistance,
                    const std::string &canonicalFwd, std::vector<Pattern> &out, std::string &err) {
    std::string canonicalRev;
    if (!reverseComplement(canonicalFwd, canonicalRev)) {
        err = "invalid nucleotide in canonical pattern '" + canonicalFwd + "'";
        return false;
    }
    std::vector<std::pair<std::string, bool>> all;
    for (const std::string &seed : rawPatterns) {
        if (seed.empty()) continue;
        for (const std::string &combo : iupacExpand(seed)) {
            // orientation of the seed: closer (Hamming, ties -> forward) to canonicalFwd or Rev
            bool forward;
            if (combo.size() == canonicalFwd.size()) {
                unsigned df = 0, dr = 0;
                for (size_t i = 0; i < combo.size(); ++i) {
                    df += combo[i] != canonicalFwd[i];
                    dr += combo[i] != canonicalRev[i];
                }
                forward = (df & 0xFFu) <= (dr & 0xFFu);
            } else {
                std::string comboRc;
                reverseComplement(combo, comboRc);
                const bool comboShorter = combo.size() < canonicalFwd.size();
                const std::string &shorter = comboShorter ? combo : canonicalFwd;
                forward = hammingAtBestOffset(shorter, comboShorter ? canonicalFwd : combo) <=
                          hammingAtBestOffset(shorter, comboShorter ? canonicalRev : comboRc);
            }
            std::vector<std::string> variants{combo};
            if (editDistance >= 1) appendNeighbours(combo, variants);
            if (editDistance >= 2) {
                const size_t d1End = variants.size();
                for (size_t v = 1; v < d1End; ++v) {
                    const std::string src = variants[v];   // a copy: appendNeighbours grows (reallocates) `variants`
                    appendNeighbours(src, variants);
                }
            }
            for (const std::string &v : variants) {
                std::string rc;
                if (!reverseComplement(v, rc)) {
                    err = "invalid nucleotide in pattern '" + v + "'";
                    return false;
                }
                all.emplace_back(v, forward);
                all.emplace_back(std::move(rc), !forward);
            }
        }
    }
    // Same std::sort (by string only) + std::unique (first survivor) over the same element order
    // as the reference, so that a string produced with both flags resolves identically.
    std::sort(all.begin(), all.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    all.erase(std::unique(all.begin(), all.end(),
                          [](const auto &a, const auto &b) { return a.first == b.first; }),
              all.end());
    out.clear();
    out.reserve(all.size());
    for (auto &p : all) {
        Pattern q;
        q.isForward = p.second;
        q.isCanonical = p.first == canonicalFwd || p.first == canonicalRev;
        q.text = std::move(p.first);
        out.push_back(std::move(q));
    }
    return true;
}

bool buildScanTables(const std::vector<Pattern> &patterns, ScanTables &t, std::string &err) {
    if (patterns.empty()) { err = "empty pattern table"; return false; }
    t = ScanTables();
    t.itsMinLen = static_cast<uint16_t>(2 * patterns.front().text.size());

    // The Trie skips non-ACGT letters on insert (include/teloscope.h:44) and stores one flag set
    // per leaf, later inserts overwriting earlier ones; reproduce both on the effective strings.
    struct Eff { std::string s; uint8_t flags; };
    std::vector<Eff> eff;
    size_t longestRaw = 0;
    for (const Pattern &p : patterns) {
        std::string s;
        for (char c : p.text) if (c == 'A' || c == 'C' || c == 'G' || c == 'T') s.push_back(c);
        longestRaw = std::max(longestRaw, p.text.size());
        if (s.empty()) continue;                       // would mark the root; can never be reported
        if (s.size() > static_cast<size_t>(kMaxPatternLen)) {
            err = "pattern longer than " + std::to_string(kMaxPatternLen) + " bases";
            return false;
        }
        uint8_t f = 1u | (p.isForward ? 2u : 0u) | (p.isCanonical ? 4u : 0u);
        auto it = std::find_if(eff.begin(), eff.end(), [&](const Eff &e) { return e.s == s; });
        if (it != eff.end()) it->flags = f; else eff.push_back({s, f});
    }
    if (eff.empty()) { err = "no ACGT pattern"; return false; }

    std::vector<uint32_t> lens;
    for (const Eff &e : eff) lens.push_back(static_cast<uint32_t>(e.s.size()));
    std::sort(lens.begin(), lens.end());
    lens.erase(std::unique(lens.begin(), lens.end()), lens.end());
    if (lens.size() > static_cast<size_t>(kMaxLenClasses)) {
        err = "more than " + std::to_string(kMaxLenClasses) + " distinct pattern lengths";
        return false;
    }
    t.minLen = lens.front();
    t.maxLen = lens.back();
    // longestPatternSize bounds the trie walk (src/teloscope.cpp:439); with only ACGT patterns it
    // equals maxLen.  A longer raw (non-ACGT-containing) pattern cannot occur after expansion.
    (void)longestRaw;
    int np = kIndexBases + 1 - static_cast<int>(t.minLen);
    t.npos = static_cast<uint32_t>(std::max(1, std::min(4, np)));

    auto pack = [](const std::string &s, size_t n) {
        uint64_t k = 0;
        for (size_t j = 0; j < n; ++j) k |= static_cast<uint64_t>(baseCode(s[j])) << (2 * j);
        return k;
    };

    t.index8.assign(1u << (2 * kIndexBases), 0);
    for (const Eff &e : eff) {
        for (uint32_t o = 0; o < t.npos; ++o) {
            const uint32_t fixed = std::min<uint32_t>(static_cast<uint32_t>(e.s.size()), kIndexBases - o);
            const uint32_t fixedKey = static_cast<uint32_t>(pack(e.s, fixed)) << (2 * o);
            const uint32_t fixedMask = ((fixed == 16 ? 0u : (1u << (2 * fixed))) - 1u) << (2 * o);
            // enumerate the free bits (everything outside fixedMask within 16 bits)
            const uint32_t freeMask = 0xFFFFu & ~fixedMask;
            uint32_t sub = 0;
            do {
                t.index8[fixedKey | sub] |= static_cast<uint8_t>(1u << o);
                sub = (sub - freeMask) & freeMask;     // next subset of freeMask
            } while (sub != 0);
        }
    }

    for (uint32_t len : lens) {
        LenClass c;
        c.len = len;
        c.direct = len <= static_cast<uint32_t>(kMaxDirectLen);
        c.sortedBegin = static_cast<uint32_t>(t.sortedKeys.size());
        std::vector<std::pair<uint64_t, uint8_t>> ks;
        for (const Eff &e : eff) if (e.s.size() == len) ks.emplace_back(pack(e.s, len), e.flags);
        std::sort(ks.begin(), ks.end());
        for (auto &k : ks) { t.sortedKeys.push_back(k.first); t.sortedFlags.push_back(k.second); }
        c.sortedCount = static_cast<uint32_t>(ks.size());
        if (c.direct) {
            c.flagOffset = static_cast<uint32_t>(t.flagTables.size());
            t.flagTables.resize(t.flagTables.size() + (size_t(1) << (2 * len)), 0);
            for (auto &k : ks) t.flagTables[c.flagOffset + k.first] = k.second;
        }
        t.classes.push_back(c);
    }
    return true;
}

}  // namespace tls
