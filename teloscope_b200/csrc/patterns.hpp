// Host-side pattern table and the device lookup tables derived from it.
//
// PatternTable mirrors what the reference keeps in userInput.patternInfo plus the isCanonical
// flag the Trie stores per leaf (reference: src/tools.cpp:201-283, include/teloscope.h:241-247).
// ScanTables is what replaces the Trie walk on the device: a k-mer -> pattern-bitmask lookup.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace tls {

struct Pattern {
    std::string text;     // ACGT only
    bool isForward = false;
    bool isCanonical = false;
};

// expandPatternsWithOrientation: IUPAC expansion, <=x substitutions, reverse complements,
// orientation by distance to canonicalFwd/Rev, sort + unique.  Returns false (err set) on an
// invalid nucleotide, for which the reference's revCom throws.
bool expandPatterns(const std::vector<std::string> &rawPatterns, int editDistance,
                    const std::string &canonicalFwd, std::vector<Pattern> &out, std::string &err);

constexpr int kIndexBases = 8;          // bases per stage-1 lookup (16-bit index)
constexpr int kMaxLenClasses = 8;       // distinct pattern lengths supported
constexpr int kMaxDirectLen = 7;        // 4^len-byte flag tables live in shared memory up to here
constexpr int kMaxPatternLen = 32;      // 2-bit packed into one u64

// Base code: (ascii >> 1) & 3  ->  A=0, C=1, T=2, G=3
inline uint32_t baseCode(char c) { return (static_cast<uint32_t>(c) >> 1) & 3u; }

struct LenClass {
    uint32_t len = 0;
    uint32_t flagOffset = 0;            // byte offset of this class's direct table in flagTables
    bool direct = false;                // len <= kMaxDirectLen
    uint32_t sortedBegin = 0, sortedCount = 0;   // slice of sortedKeys/sortedFlags (all classes)
};

struct ScanTables {
    // stage 1: 65536 bytes, bit o (o < npos) set when some pattern is compatible with the
    // 8-mer at offset o (exact for patterns with o+len <= 8, prefix test otherwise)
    std::vector<uint8_t> index8;
    uint32_t npos = 1;                  // positions resolved per lookup: clamp(9 - minLen, 1, 4)
    // stage 2: per length class
    std::vector<LenClass> classes;      // ascending length
    std::vector<uint8_t> flagTables;    // direct tables back to back; entry: bit0 hit, bit1 fwd, bit2 canonical
    std::vector<uint64_t> sortedKeys;   // little-endian 2-bit packed patterns, sorted per class
    std::vector<uint8_t> sortedFlags;
    uint32_t minLen = 0, maxLen = 0;
    uint32_t itsMinLen = 0;             // 2 * patterns.front().size(), src/teloscope.cpp:185
};

bool buildScanTables(const std::vector<Pattern> &patterns, ScanTables &out, std::string &err);

}  // namespace tls
