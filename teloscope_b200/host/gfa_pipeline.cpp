// teloscope-b200 on a GFA assembly graph: telomere annotation of the graph.
//
// What the reference does for a .gfa input (Input::read, src/input.cpp:625-716; walkSegment / walkSegmentForPath,
// :835-939; gfalibs reader gfalibs/src/input-gfa.cpp:150-794 and GFA1.2 writer gfalibs/src/output.cpp:339-447):
//   * the graph is read (GFA1 S / L / J / P lines; GFA2 S lines), kept in input order;
//   * one scan job per TERMINAL END: with paths, the first and the last segment of every path (de-duplicated by segment,
//     path orientation and end, in the order of the reference's std::set of (segment id, orientation, isFirst)); without
//     paths, every segment;
//   * every job is a tips scan of the whole segment (scanSegment(seq, 0, tipsOnly = true)) — here: all segments that
//     need one go to the GPU as ONE TLS_MODE_TIPS batch through the C ABI;
//   * a terminal block at the path-terminal physical end of its segment (or, without paths, any terminal block) becomes a
//     synthetic node "telomere_<segment><orient>_<start|end>" (S line, sequence *, tags LN:i:6 RC:i:6000 TL:i:<block length>,
//     the longest block when several share a node) tied to the segment by an L line with a 0M overlap and RC:i:0;
//   * the graph is written back as GFA 1.2 (H, S, L, J, P in that order, synthetic nodes and links behind the input's) to
//     <out>/<input name>.telo.annotated.gfa, the node names with the colour #008000 to <...>.telo.annotated.colors.csv.
// The reference appends synthetic nodes in the order its worker threads finish; this binary always uses the order of
// -j 1 (jobs one after the other), which is what the reference's own tests pin.
#include "pipeline.hpp"

#include <zlib.h>

#include <map>
#include <set>
#include <tuple>
#include <unordered_map>

namespace {

using namespace tlsb;

struct Tag { std::string label, content; char type; };          // gfalibs Tag: two-letter label, one-letter type, content
struct Segment { std::string name, seq; std::vector<Tag> tags; uint32_t id; bool hasLine; };
struct Edge { uint32_t a, b; char oa, ob; std::string cigar; std::vector<Tag> tags; };
struct GapRec { uint32_t a, b; char oa, ob; long dist; std::vector<Tag> tags; };
struct Comp { int type; uint32_t id; char orient; unsigned start, end; };      // type 0 segment, 1 edge (','), 2 gap (';')
struct PathRec { std::string name, comment; std::vector<Comp> comps; };

std::vector<std::string> split(const std::string &s, char d) {     // gfalibs readDelimited: pieces between delimiters, a trailing piece if non-empty
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t b = s.find(d, a);
        if (b == std::string::npos) { if (a < s.size()) out.push_back(s.substr(a)); break; }
        out.push_back(s.substr(a, b - a));
        a = b + 1;
    }
    return out;
}

std::vector<Tag> parseTags(const std::vector<std::string> &f, size_t from) {
    std::vector<Tag> tags;
    for (size_t i = from; i < f.size(); ++i) {
        const std::vector<std::string> v = split(f[i], ':');
        if (v.size() < 3 || v[0].size() < 2 || v[1].empty()) die("Error: malformed GFA tag '%s'.\n", f[i].c_str());
        tags.push_back(Tag{v[0].substr(0, 2), v[2], v[1][0]});
    }
    return tags;
}

void writeTags(std::string &out, const std::vector<Tag> &tags) {
    for (const Tag &t : tags) { out += '\t'; out += t.label; out += ':'; out += t.type; out += ':'; out += t.content; }
}

struct Graph {
    std::vector<Segment> segments;                         // S lines in input order, then the synthetic nodes
    std::unordered_map<std::string, uint32_t> byName;      // header -> index in `names`
    std::vector<std::string> names;                        // every header the file mentions, in order of first mention
    std::vector<int> segOfName;                            // index into segments, -1: mentioned but without an S line
    std::vector<Edge> edges;
    std::vector<GapRec> gaps;
    std::vector<PathRec> paths;

    uint32_t touch(const std::string &name) {
        auto it = byName.find(name);
        if (it != byName.end()) return it->second;
        const uint32_t id = static_cast<uint32_t>(names.size());
        byName.emplace(name, id);
        names.push_back(name);
        segOfName.push_back(-1);
        return id;
    }
    void addSegment(const std::string &name, std::string seq, std::vector<Tag> tags) {
        const uint32_t id = touch(name);
        if (segOfName[id] >= 0) die("Error: segment name already exists (%s). Terminating.\n", name.c_str());
        segOfName[id] = static_cast<int>(segments.size());
        segments.push_back(Segment{name, std::move(seq), std::move(tags), id, true});
    }
    bool linked(uint32_t a, uint32_t b) const {
        for (const Edge &e : edges) if ((e.a == a && e.b == b) || (e.a == b && e.b == a)) return true;
        return false;
    }
    bool gapped(uint32_t a, uint32_t b) const {
        for (const GapRec &g : gaps) if ((g.a == a && g.b == b) || (g.a == b && g.b == a)) return true;
        return false;
    }
};

// P line, column 3: components with ',' (link) or ';' (gap) between them, an optional leading / trailing separator for
// a terminal gap, an optional (start:end) range before the orientation (gfalibs/src/input-gfa.cpp:643-780).  An EDGE /
// GAP component only enters the path when the graph holds that link / gap at this point of the file.
void parsePath(Graph &g, const std::vector<std::string> &f) {
    PathRec p;
    p.name = f[1];
    if (g.byName.count(p.name)) die("Error: path name already exists (%s). Terminating.\n", p.name.c_str());
    g.touch(p.name);
    const std::string &txt = f.size() > 2 ? f[2] : std::string();
    // tokens: text up to and including each separator
    std::vector<std::string> tok;
    {
        std::string cur;
        for (char ch : txt) { cur += ch; if (ch == ',' || ch == ';') { tok.push_back(cur); cur.clear(); } }
        if (!cur.empty()) tok.push_back(cur);
    }
    auto parseComp = [&](std::string c, char &orient, unsigned &start, unsigned &end) {
        if (c.empty()) die("Error: empty component in path (%s). Terminating.\n", p.name.c_str());
        orient = c.back();
        c.pop_back();
        start = end = 0;
        const size_t lp = c.find('('), colon = c.find(':'), rp = c.find(')');
        if (lp != std::string::npos && colon != std::string::npos && rp != std::string::npos) {
            start = static_cast<unsigned>(std::stoi(c.substr(lp + 1, colon - lp - 1)));
            end = static_cast<unsigned>(std::stoi(c.substr(colon + 1, rp - colon - 1)));
            c = c.substr(0, lp);
        }
        return c;
    };
    for (size_t i = 0; i < tok.size(); ++i) {
        std::string t = tok[i];
        char sep = 0;
        if (!t.empty() && (t.back() == ',' || t.back() == ';')) { sep = t.back(); t.pop_back(); }
        if (t.empty()) {                                   // a leading separator: terminal gap before the first segment
            if (i + 1 < tok.size() && sep == ';') {
                std::string nx = tok[i + 1];
                if (!nx.empty() && (nx.back() == ',' || nx.back() == ';')) nx.pop_back();
                char o; unsigned s, e;
                const std::string name = parseComp(nx, o, s, e);
                auto it = g.byName.find(name);
                if (it == g.byName.end()) die("Error: cannot find next component in path (%s). Terminating.\n", name.c_str());
                if (g.gapped(it->second, it->second)) p.comps.push_back(Comp{2, 0, '0', 0, 0});
            }
            continue;
        }
        char o; unsigned s, e;
        const std::string name = parseComp(t, o, s, e);
        const uint32_t id = g.touch(name);
        p.comps.push_back(Comp{0, id, o, s, e});
        if (sep == 0) continue;
        if (i + 1 < tok.size()) {
            std::string nx = tok[i + 1];
            if (!nx.empty() && (nx.back() == ',' || nx.back() == ';')) nx.pop_back();
            char o2; unsigned s2, e2;
            const std::string name2 = parseComp(nx, o2, s2, e2);
            auto it = g.byName.find(name2);
            if (it == g.byName.end()) die("Error: cannot find next component in path (%s). Terminating.\n", name2.c_str());
            if (sep == ',' ? g.linked(id, it->second) : g.gapped(id, it->second)) p.comps.push_back(Comp{sep == ',' ? 1 : 2, 0, '0', s, e});
        } else if (g.gapped(id, id)) {                     // a trailing separator: terminal gap behind the last segment
            p.comps.push_back(Comp{2, 0, '0', s, e});
        }
    }
    for (size_t i = 2; i < f.size(); ++i)
        if (f[i].compare(0, 3, "C:Z") == 0) p.comment = f[i].substr(std::min<size_t>(4, f[i].size()));
    g.paths.push_back(std::move(p));
}

void readGraph(const InputData &in, Graph &g) {
    size_t pos = 0;
    auto nextLine = [&](std::string &line) {
        if (pos >= in.n) return false;
        const char *nl = static_cast<const char *>(memchr(in.p + pos, '\n', in.n - pos));
        const size_t e = nl ? static_cast<size_t>(nl - in.p) : in.n;
        line.assign(in.p + pos, e - pos);
        pos = nl ? e + 1 : in.n;
        return true;
    };
    std::string line, version = "1";
    bool first = true;
    while (nextLine(line)) {
        if (line.empty()) { first = false; continue; }
        const std::vector<std::string> f = split(line, '\t');
        if (first && line[0] == 'H') {                     // H <tab> VN:Z:<version>
            if (f.size() > 1) { const std::vector<std::string> v = split(f[1], ':'); if (v.size() > 2 && !v[2].empty()) version = v[2]; }
            first = false;
            continue;
        }
        if (first && (line[0] == 'G' || line[0] == 'O' || line[0] == 'E')) version = "2";
        first = false;
        if (version[0] == '2') {
            switch (line[0]) {
                case 'S':                                  // S name length sequence tags
                    if (f.size() < 4) die("Error: malformed GFA2 segment line.\n");
                    g.addSegment(f[1], f[3], parseTags(f, 4));
                    break;
                case 'E': case 'G': case 'O': case 'U': case 'F':
                    unsupported("GFA2 edge / gap / group lines");
                default: break;
            }
            continue;
        }
        switch (line[0]) {
            case 'S':
                if (f.size() < 3) die("Error: malformed GFA segment line.\n");
                g.addSegment(f[1], f[2], parseTags(f, 3));
                break;
            case 'L': {
                if (f.size() < 6) die("Error: malformed GFA link line.\n");
                Edge e;
                e.a = g.touch(f[1]); e.oa = f[2].empty() ? '+' : f[2][0];
                e.b = g.touch(f[3]); e.ob = f[4].empty() ? '+' : f[4][0];
                e.cigar = f[5];
                if (f.size() > 6 && !f[6].empty()) e.tags = parseTags(f, 6);
                g.edges.push_back(std::move(e));
                break;
            }
            case 'J': {
                if (f.size() < 6) die("Error: malformed GFA jump line.\n");
                GapRec j;
                j.a = g.touch(f[1]); j.oa = f[2].empty() ? '+' : f[2][0];
                j.b = g.touch(f[3]); j.ob = f[4].empty() ? '+' : f[4][0];
                try { j.dist = std::stoi(f[5]); } catch (const std::exception &) { die("Error: malformed GFA jump distance '%s'.\n", f[5].c_str()); }
                j.tags = parseTags(f, 6);
                g.gaps.push_back(std::move(j));
                break;
            }
            case 'P':
                if (f.size() < 3) die("Error: malformed GFA path line.\n");
                parsePath(g, f);
                break;
            default: break;                                // comments and line types gfalibs does not keep
        }
    }
}

void writeGraph(const Graph &g, const std::string &path) {
    std::string out = "H\tVN:Z:1.2\n";
    for (const Segment &s : g.segments) {
        out += "S\t"; out += s.name; out += '\t'; out += s.seq;
        writeTags(out, s.tags);
        out += '\n';
    }
    for (const Edge &e : g.edges) {
        out += "L\t"; out += g.names[e.a]; out += '\t'; out += e.oa; out += '\t'; out += g.names[e.b]; out += '\t'; out += e.ob; out += '\t'; out += e.cigar;
        writeTags(out, e.tags);
        out += '\n';
    }
    for (const GapRec &j : g.gaps) {
        out += "J\t"; out += g.names[j.a]; out += '\t'; out += j.oa; out += '\t'; out += g.names[j.b]; out += '\t'; out += j.ob; out += '\t';
        out += j.dist != 0 ? std::to_string(j.dist) : std::string("*");
        writeTags(out, j.tags);
        out += '\n';
    }
    for (const PathRec &p : g.paths) {
        out += "P\t"; out += p.name; out += '\t';
        for (const Comp &c : p.comps) {
            if (c.type == 0) {
                out += g.names[c.id];
                if (c.start != 0 || c.end != 0) { out += '('; out += std::to_string(c.start); out += ':'; out += std::to_string(c.end); out += ')'; }
                out += c.orient;
            } else {
                out += c.type == 1 ? ',' : ';';
            }
        }
        out += "\t*";
        if (!p.comment.empty()) { out += "\tCM:Z:"; out += p.comment; }
        out += '\n';
    }
    FILE *f = fopen(path.c_str(), "wb");
    if (!f || fwrite(out.data(), 1, out.size(), f) != out.size()) die("Error: Cannot write annotated GFA to '%s'.\n", path.c_str());
    fclose(f);
}

}  // namespace

namespace tlsb {

bool isGfaPath(const std::string &path) {                  // isGfaAssemblyPath: .gfa / .gfa.gz / .gfa2 ...
    std::string lower = path;
    for (char &c : lower) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    auto ends = [&](const char *suf) { const size_t n = strlen(suf); return lower.size() >= n && lower.compare(lower.size() - n, n, suf) == 0; };
    return ends(".gfa") || ends(".gfa.gz") || ends(".gfa2") || ends(".gfa2.gz");
}

int normalizeGfa(const std::string &inPath, const std::string &outPath) {
    InputData in;
    loadInput(inPath, in);
    Graph g;
    readGraph(in, g);
    writeGraph(g, outPath);
    return 0;
}

int runGfa(const Options &o, const tls_config &baseCfg) {
    StageTimer timer;
    tls_config cfg = baseCfg;
    cfg.outputs = 0;
    tls_ctx *ctx = nullptr;
    int ctxRc = TLS_OK;
    std::string ctxErr;
    std::thread ctxThread([&]() {                          // the CUDA context comes up while the graph is read
        const char *devEnv = getenv("TLS_DEVICE");
        ctxRc = tls_create(&cfg, devEnv ? atoi(devEnv) : 0, &ctx);
        if (ctxRc != TLS_OK) ctxErr = tls_last_error();
    });
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{ctxThread};

    InputData in;
    loadInput(o.inSequence, in);
    Graph g;
    readGraph(in, g);
    timer.mark("read graph");

    // ---- terminal ends (src/input.cpp:634-671) ------------------------------------------------------------------------
    struct Job { uint32_t nameId; char orient; bool isFirst, pathAware; };
    std::vector<Job> jobs;
    if (!g.paths.empty()) {
        std::set<std::tuple<uint32_t, char, bool>> ends;
        for (const PathRec &p : g.paths) {
            for (const Comp &c : p.comps) if (c.type == 0 && c.orient != '0') { ends.emplace(c.id, c.orient, true); break; }
            for (auto it = p.comps.rbegin(); it != p.comps.rend(); ++it) if (it->type == 0 && it->orient != '0') { ends.emplace(it->id, it->orient, false); break; }
        }
        for (const auto &e : ends) jobs.push_back(Job{std::get<0>(e), std::get<1>(e), std::get<2>(e), true});
    } else {
        for (const Segment &s : g.segments) jobs.push_back(Job{s.id, '+', false, false});
    }
    size_t noSeq = 0;
    for (const Job &j : jobs) {
        const int si = g.segOfName[j.nameId];
        if (si < 0 || g.segments[static_cast<size_t>(si)].seq == "*") ++noSeq;
    }
    if (noSeq > 0)
        fprintf(stderr, "Warning: %zu of %zu GFA segment(s) had no sequence (*); skipped for telomere annotation.\n", noSeq, jobs.size());

    // ---- one tips batch with every segment that has a job -----------------------------------------------------------
    std::vector<int> itemOfSeg(g.segments.size(), -1);
    std::vector<uint64_t> offsets{0};
    std::string seq;
    std::vector<uint32_t> segOfItem;
    for (const Job &j : jobs) {
        const int si = g.segOfName[j.nameId];
        if (si < 0 || g.segments[static_cast<size_t>(si)].seq == "*" || itemOfSeg[static_cast<size_t>(si)] >= 0) continue;
        itemOfSeg[static_cast<size_t>(si)] = static_cast<int>(segOfItem.size());
        segOfItem.push_back(static_cast<uint32_t>(si));
        seq += g.segments[static_cast<size_t>(si)].seq;
        offsets.push_back(seq.size());
    }
    tls_result res{};
    int ticket = -1;
    if (!segOfItem.empty()) {
        ctxThread.join();
        if (ctxRc != TLS_OK) die("Error: tls_create failed: %s\n", ctxErr.c_str());
        seq.resize(seq.size() + 256, '\0');
        tls_batch tb{};
        tb.seq = reinterpret_cast<const uint8_t *>(seq.data()); tb.offsets = offsets.data(); tb.n_items = static_cast<uint32_t>(segOfItem.size());
        tb.mode = TLS_MODE_TIPS;
        check(tls_submit(ctx, &tb, &ticket), "tls_submit");
        check(tls_collect(ctx, ticket, &res), "tls_collect");
    }
    timer.mark("scan");

    // ---- synthetic nodes and links, job by job (walkSegment / walkSegmentForPath) --------------------------------------
    const size_t nInputSegments = g.segments.size();
    for (const Job &j : jobs) {
        const int si = g.segOfName[j.nameId];
        if (si < 0 || itemOfSeg[static_cast<size_t>(si)] < 0) continue;
        const Segment &seg = g.segments[static_cast<size_t>(si)];
        const size_t item = static_cast<size_t>(itemOfSeg[static_cast<size_t>(si)]);
        const uint64_t n = seg.seq.size();
        struct Ann { std::string header; uint32_t blockLen; char segOrient; };
        std::vector<Ann> anns;
        const bool scanStart = j.isFirst == (j.orient == '+');
        for (uint64_t b = res.terminal_off[item]; b < res.terminal_off[item + 1]; ++b) {
            const tls_block &blk = res.terminal[b];
            const uint64_t distToStart = blk.start, distToEnd = n - (blk.start + blk.len);
            const bool atStart = distToStart <= distToEnd;
            std::string header;
            char segOrient;
            if (j.pathAware) {
                if (atStart != scanStart) continue;
                header = "telomere_" + seg.name + std::string(1, j.orient) + "_" + (j.isFirst ? "start" : "end");
                segOrient = j.isFirst ? j.orient : (j.orient == '+' ? '-' : '+');
            } else {
                header = "telomere_" + seg.name + "+" + (atStart ? "_start" : "_end");
                segOrient = atStart ? '+' : '-';
            }
            bool merged = false;
            for (Ann &a : anns) if (a.header == header) { a.blockLen = std::max(a.blockLen, blk.len); merged = true; break; }
            if (!merged) anns.push_back(Ann{header, blk.len, segOrient});
        }
        for (const Ann &a : anns) {
            if (g.byName.count(a.header)) continue;       // terminal ends are unique; a clash with an input name keeps the input's
            std::vector<Tag> tags{Tag{"LN", "6", 'i'}, Tag{"RC", "6000", 'i'}, Tag{"TL", std::to_string(a.blockLen), 'i'}};
            g.addSegment(a.header, "*", std::move(tags));
            Edge e;
            e.a = g.byName.at(a.header); e.oa = '+'; e.b = j.nameId; e.ob = a.segOrient; e.cigar = "0M";
            e.tags.push_back(Tag{"RC", "0", 'i'});
            g.edges.push_back(std::move(e));
        }
    }
    if (ticket >= 0) check(tls_release(ctx, ticket), "tls_release");

    // ---- outputs ------------------------------------------------------------------------------------------------------
    writeGraph(g, o.outRoute + "/" + o.inSequenceName + ".telo.annotated.gfa");
    const std::string colors = o.outRoute + "/" + o.inSequenceName + ".telo.annotated.colors.csv";
    if (FILE *f = fopen(colors.c_str(), "wb")) {
        std::string c = "node\tcolor\n";
        for (size_t i = 0; i < g.segments.size(); ++i)
            if (g.segments[i].name.rfind("telomere_", 0) == 0) { c += g.segments[i].name; c += "\t#008000\n"; }
        fwrite(c.data(), 1, c.size(), f);
        fclose(f);
    } else {
        fprintf(stderr, "Warning: could not write telomere colours CSV to %s.\n", colors.c_str());
    }
    (void)nInputSegments;
    timer.mark("write");
    fflush(stdout);
    fflush(stderr);
    _exit(EXIT_SUCCESS);
}

}  // namespace tlsb
