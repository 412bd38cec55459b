// bvg_format.hpp -- host side of the loader: .properties parsing, compression flags, and the one sequential
// pass over the .offsets stream (gamma/delta coded gaps -> absolute bit offsets).
//
// Mirrors BVGraph.loadInternal (reference src/it/unimi/dsi/webgraph/BVGraph.java:1516-1609), setFlags/string2Flags
// (:1317-1366) and OffsetsLongIterator (:907-935).  This is the host half of the drop-in boundary; it never touches
// successor lists (those are decoded on the GPU only).
#pragma once
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace bvg {

struct Properties {
    int64_t nodes = -1, arcs = -1;
    int32_t window = -1, maxref = -1, minlen = -1, zetak = 3;
    uint32_t flags = 0;
};

inline int coding_id(const std::string& s) {
    static const char* names[] = { "", "DELTA", "GAMMA", "GOLOMB", "SKEWED_GOLOMB", "UNARY", "ZETA", "NIBBLE" };
    for (int i = 1; i < 8; i++) if (s == names[i]) return i;
    return -1;
}

// "OUTDEGREES_DELTA | RESIDUALS_GAMMA" -> flag word (BVGraph.java:1352-1366, constants :474-523). -1 on unknown names.
inline int64_t parse_flags(const std::string& str) {
    static const struct { const char* prefix; int shift; } slots[] = {
        { "OUTDEGREES_", 0 }, { "BLOCK_COUNT_", 16 }, { "BLOCKS_", 4 }, { "RESIDUALS_", 8 }, { "REFERENCES_", 12 }, { "OFFSETS_", 20 } };
    uint32_t flags = 0;
    size_t p = 0;
    while (p <= str.size()) {
        size_t q = str.find('|', p);
        if (q == std::string::npos) q = str.size();
        std::string tok = str.substr(p, q - p);
        p = q + 1;
        size_t a = 0, b = tok.size();
        while (a < b && isspace((unsigned char)tok[a])) a++;
        while (b > a && isspace((unsigned char)tok[b - 1])) b--;
        tok = tok.substr(a, b - a);
        if (tok.empty()) continue;
        bool found = false;
        for (const auto& s : slots) {
            const size_t pl = strlen(s.prefix);
            if (tok.compare(0, pl, s.prefix) == 0) {
                const int c = coding_id(tok.substr(pl));
                if (c < 0) return -1;
                flags |= (uint32_t)c << s.shift;
                found = true;
                break;
            }
        }
        if (!found) return -1;
    }
    return flags;
}

// Minimal java.util.Properties reader (key=value | key:value | key value, '#'/'!' comments).
inline bool read_properties_file(const std::string& path, std::map<std::string, std::string>& kv) {
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096];
    while (fgets(line, sizeof line, f)) {
        char* s = line;
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s || *s == '#' || *s == '!') continue;
        char* e = s;
        while (*e && *e != '=' && *e != ':' && !isspace((unsigned char)*e)) e++;
        std::string key(s, e - s);
        while (*e && (isspace((unsigned char)*e))) e++;
        if (*e == '=' || *e == ':') e++;
        while (*e && (*e == ' ' || *e == '\t')) e++;
        std::string val(e);
        while (!val.empty() && isspace((unsigned char)val.back())) val.pop_back();
        kv[key] = val;
    }
    fclose(f);
    return true;
}

// Returns 0, or -4 (EIO) / -5 (EFORMAT) / -1 (EINVAL) following loadInternal's checks.
inline int load_properties(const std::string& basename, Properties& p) {
    std::map<std::string, std::string> kv;
    if (!read_properties_file(basename + ".properties", kv)) return -4;
    auto has = [&](const char* k) { return kv.find(k) != kv.end(); };
    if (!has("graphclass")) return -5;
    const std::string gc = kv["graphclass"];  // :1528, the "big" package is accepted too
    if (gc != "it.unimi.dsi.webgraph.BVGraph" && gc != "it.unimi.dsi.big.webgraph.BVGraph") return -5;
    if (!has("version") || atoi(kv["version"].c_str()) > 0) return -5;  // :1533-1534
    const int64_t fl = parse_flags(has("compressionflags") ? kv["compressionflags"] : std::string());
    if (fl < 0) return -5;
    p.flags = (uint32_t)fl;
    if (!has("nodes") || !has("arcs") || !has("windowsize") || !has("maxrefcount") || !has("minintervallength")) return -5;
    p.nodes = atoll(kv["nodes"].c_str());
    if (p.nodes > 2147483647LL || p.nodes < 0) return -1;  // :1537
    p.arcs = atoll(kv["arcs"].c_str());
    p.window = atoi(kv["windowsize"].c_str());
    p.maxref = atoi(kv["maxrefcount"].c_str());
    p.minlen = atoi(kv["minintervallength"].c_str());
    if (has("zetak")) p.zetak = atoi(kv["zetak"].c_str());
    if (p.window < 0 || p.minlen < 0 || p.zetak < 1) return -5;
    return 0;
}

// MSB-first reader for the .offsets stream only.
struct HostBits {
    const uint8_t* buf;  // >= 16 readable bytes past the end
    uint64_t nbits, pos = 0;
    inline uint64_t peek() const {
        uint64_t w;
        memcpy(&w, buf + (pos >> 3), 8);
        return __builtin_bswap64(w) << (pos & 7);
    }
    inline uint64_t bits(int n) {
        uint64_t r = 0;
        while (n > 32) { r = (r << 32) | (peek() >> 32); pos += 32; n -= 32; }
        if (n > 0) { r = (r << n) | (peek() >> (64 - n)); pos += (uint64_t)n; }
        return r;
    }
    inline uint64_t unary() {
        uint64_t zeros = 0;
        for (;;) {
            const int valid = 64 - (int)(pos & 7);
            const uint64_t w = peek();
            if (w == 0) { zeros += (uint64_t)valid; pos += (uint64_t)valid; if (pos > nbits + 64) return zeros; continue; }
            const int z = __builtin_clzll(w);
            pos += (uint64_t)z + 1;
            return zeros + (uint64_t)z;
        }
    }
    inline uint64_t gamma() { const uint64_t m = unary(); return m > 63 ? ~0ull : ((1ull << m) | bits((int)m)) - 1; }
    inline uint64_t delta() { const uint64_t m = gamma(); return m > 63 ? ~0ull : ((1ull << m) | bits((int)m)) - 1; }
};

// n+1 gaps -> absolute offsets (OffsetsLongIterator.nextLong, :926-934; readOffset :631-637). 0 or -4/-3.
inline int decode_offsets_stream(const uint8_t* stream, uint64_t nbytes, int offset_coding, int64_t n, std::vector<uint64_t>& out) {
    if (offset_coding != 2 /*GAMMA*/ && offset_coding != 1 /*DELTA*/) return -3;
    std::vector<uint8_t> padded(nbytes + 16, 0);
    memcpy(padded.data(), stream, nbytes);
    HostBits b{ padded.data(), nbytes * 8 };
    out.resize((size_t)n + 1);
    uint64_t off = 0;
    for (int64_t i = 0; i <= n; i++) {
        off += offset_coding == 2 ? b.gamma() : b.delta();
        out[(size_t)i] = off;
        if (b.pos > b.nbits) return -4;
    }
    return 0;
}

inline bool slurp_file(const std::string& path, std::vector<uint8_t>& out, uint64_t from = 0, uint64_t len = ~0ull) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseeko(f, 0, SEEK_END);
    const uint64_t size = (uint64_t)ftello(f);
    if (from > size) from = size;
    if (len > size - from) len = size - from;
    out.resize((size_t)len);
    fseeko(f, (off_t)from, SEEK_SET);
    const bool ok = len == 0 || fread(out.data(), 1, (size_t)len, f) == len;
    fclose(f);
    return ok;
}

inline uint64_t file_size(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return ~0ull;
    fseeko(f, 0, SEEK_END);
    const uint64_t size = (uint64_t)ftello(f);
    fclose(f);
    return size;
}

}  // namespace bvg
