/*
 * bvg_oracle.c -- CPU oracle for BVGraph decode.  TEST INFRASTRUCTURE ONLY (see bvg_oracle.h).
 *
 * Every function cites the reference lines it restates (paths relative to the reference root,
 * src/it/unimi/dsi/webgraph/ unless stated).  No line of the reference is copied: the Java builds
 * a lazy iterator pipeline, this materialises plain arrays.
 */
#include "bvg_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

/* ------------------------------------------------------------------------------------------
 * Bit input: restates dsiutils it.unimi.dsi.io.InputBitStream (external dependency, not in the
 * reference tree; call sites BVGraph.java:631-816, 1077-1092).  MSB-first: bit i of the stream is
 * bit (7 - i%8) of byte i/8.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const uint8_t* buf;   /* at least 16 readable bytes past nbits/8 */
    uint64_t nbits;       /* logical length */
    uint64_t pos;         /* current bit */
} ibs_t;

/* 64-bit window whose MSB is the bit at pos; only the top 64-(pos&7) >= 57 bits are stream bits. */
static inline uint64_t ibs_peek(const ibs_t* s) {
    uint64_t w;
    memcpy(&w, s->buf + (s->pos >> 3), 8);
    w = __builtin_bswap64(w);
    return w << (s->pos & 7);
}

/* InputBitStream.readInt(n)/readLong(n): n bits, MSB first. */
static inline uint64_t ibs_read_bits(ibs_t* s, int n) {
    uint64_t r = 0;
    while (n > 32) { /* split so that each piece fits the 57 valid bits */
        r = (r << 32) | (ibs_peek(s) >> 32);
        s->pos += 32;
        n -= 32;
    }
    if (n > 0) {
        r = (r << n) | (ibs_peek(s) >> (64 - n));
        s->pos += (uint64_t)n;
    }
    return r;
}

/* InputBitStream.readUnary(): number of zeros before the first one. */
static inline uint64_t ibs_read_unary(ibs_t* s) {
    uint64_t zeros = 0;
    for (;;) {
        const int valid = 64 - (int)(s->pos & 7);
        const uint64_t w = ibs_peek(s);
        if (w == 0) { /* the shifted-in low bits are zero too, so w == 0 <=> all valid bits are 0 */
            zeros += (uint64_t)valid;
            s->pos += (uint64_t)valid;
            if (s->pos > s->nbits + 64) return zeros; /* ran off the stream: caller detects */
            continue;
        }
        const int z = __builtin_clzll(w);
        s->pos += (uint64_t)z + 1;
        return zeros + (uint64_t)z;
    }
}

/* readGamma / readLongGamma: unary(msb) then msb low bits of x+1. */
static inline uint64_t ibs_read_gamma(ibs_t* s) {
    const int msb = (int)ibs_read_unary(s);
    if (msb > 63) return ~(uint64_t)0; /* corrupt; caller runs past nbits and reports */
    return (((uint64_t)1 << msb) | ibs_read_bits(s, msb)) - 1;
}

/* readDelta / readLongDelta: gamma(msb) then msb low bits of x+1. */
static inline uint64_t ibs_read_delta(ibs_t* s) {
    const uint64_t msb = ibs_read_gamma(s);
    if (msb > 63) return ~(uint64_t)0;
    return (((uint64_t)1 << msb) | ibs_read_bits(s, (int)msb)) - 1;
}

/* readZeta(k) / readLongZeta(k): unary(h); minimal binary code of x+1-2^{hk} in [0, 2^{(h+1)k}-2^{hk}). */
static inline uint64_t ibs_read_zeta(ibs_t* s, int k) {
    const uint64_t h = ibs_read_unary(s);
    if (h * (uint64_t)k + (uint64_t)k > 64) return ~(uint64_t)0;
    const uint64_t left = (uint64_t)1 << (h * (uint64_t)k);
    const uint64_t v = ibs_read_bits(s, (int)(h * (uint64_t)k) + k - 1);
    if (v < left) return v + left - 1;
    return ((v << 1) | ibs_read_bits(s, 1)) - 1;
}

/* it.unimi.dsi.bits.Fast.nat2int: even -> v/2, odd -> -(v+1)/2. */
static inline int64_t nat2int(uint64_t v) {
    return (v & 1) ? -(int64_t)((v + 1) >> 1) : (int64_t)(v >> 1);
}

/* readGolomb(b) / readLongGolomb(b) (dsiutils InputBitStream; BVGraph passes zetaK as b, BVGraph.java:796, 812):
 * b == 0 reads nothing; else unary(x / b) followed by readMinimalBinary(b): with l = msb(b), m = 2^(l+1) - b,
 * read l bits r; r < m is the remainder, otherwise one more bit is read and the remainder is 2r + bit - m.
 * Parity unpinned: no reference test or fixture uses it. */
static inline uint64_t ibs_read_golomb(ibs_t* s, int b) {
    if (b <= 0) return 0;
    const uint64_t q = ibs_read_unary(s);
    const int l = 31 - __builtin_clz((unsigned)b);
    const uint64_t m = ((uint64_t)2 << l) - (uint64_t)b;
    uint64_t r = ibs_read_bits(s, l);
    if (r >= m) r = ((r << 1) | ibs_read_bits(s, 1)) - m;
    return q * (uint64_t)b + r;
}

/* readNibble() / readLongNibble(): do { x <<= 3; stop = readBit(); x |= readInt(3); } while (!stop).
 * Parity unpinned, as above. */
static inline uint64_t ibs_read_nibble(ibs_t* s) {
    uint64_t x = 0;
    for (int i = 0; i < 22; i++) {
        const uint64_t stop = ibs_read_bits(s, 1);
        x = (x << 3) | ibs_read_bits(s, 3);
        if (stop) return x;
    }
    return ~(uint64_t)0; /* more than 64 bits of value: corrupt; caller runs past nbits and reports */
}

static inline uint64_t read_coded(ibs_t* s, int coding, int k, int* err) {
    switch (coding) {
        case BVGO_GAMMA:  return ibs_read_gamma(s);
        case BVGO_DELTA:  return ibs_read_delta(s);
        case BVGO_UNARY:  return ibs_read_unary(s);
        case BVGO_ZETA:   return ibs_read_zeta(s, k);
        case BVGO_GOLOMB: return ibs_read_golomb(s, k);
        case BVGO_NIBBLE: return ibs_read_nibble(s);
        default: *err = BVGO_EUNSUPPORTED; return 0; /* skewed Golomb: no reader in the reference either */
    }
}

uint64_t orc_read_code(const uint8_t* buf, uint64_t nbytes, uint64_t* bitpos, int coding, int k) {
    /* copy into a padded buffer so the window reads stay in bounds */
    uint8_t* tmp = (uint8_t*)calloc(nbytes + 32, 1);
    memcpy(tmp, buf, nbytes);
    ibs_t s = { tmp, nbytes * 8, *bitpos };
    int err = 0;
    uint64_t v = read_coded(&s, coding, k, &err);
    *bitpos = s.pos;
    free(tmp);
    return v;
}

/* ------------------------------------------------------------------------------------------
 * Properties and flags: BVGraph.loadInternal 1516-1545, setFlags 1317-1325, string2Flags 1352-1366,
 * constants 474-523.
 * ---------------------------------------------------------------------------------------- */
static void set_flags(orc_graph* g, uint32_t flags) {
    g->flags = flags;
    g->outdegree_coding = BVGO_GAMMA;   /* defaults, BVGraph.java:525-541 */
    g->block_coding = BVGO_GAMMA;
    g->residual_coding = BVGO_ZETA;
    g->reference_coding = BVGO_UNARY;
    g->block_count_coding = BVGO_GAMMA;
    g->offset_coding = BVGO_GAMMA;
    if (flags & 0xF) g->outdegree_coding = (int)(flags & 0xF);
    if ((flags >> 4) & 0xF) g->block_coding = (int)((flags >> 4) & 0xF);
    if ((flags >> 8) & 0xF) g->residual_coding = (int)((flags >> 8) & 0xF);
    if ((flags >> 12) & 0xF) g->reference_coding = (int)((flags >> 12) & 0xF);
    if ((flags >> 16) & 0xF) g->block_count_coding = (int)((flags >> 16) & 0xF);
    if ((flags >> 20) & 0xF) g->offset_coding = (int)((flags >> 20) & 0xF);
}

static int coding_by_name(const char* s) {
    if (!strcmp(s, "DELTA")) return BVGO_DELTA;
    if (!strcmp(s, "GAMMA")) return BVGO_GAMMA;
    if (!strcmp(s, "GOLOMB")) return BVGO_GOLOMB;
    if (!strcmp(s, "SKEWED_GOLOMB")) return BVGO_SKEWED_GOLOMB;
    if (!strcmp(s, "UNARY")) return BVGO_UNARY;
    if (!strcmp(s, "ZETA")) return BVGO_ZETA;
    if (!strcmp(s, "NIBBLE")) return BVGO_NIBBLE;
    return -1;
}

/* "OUTDEGREES_DELTA | RESIDUALS_GAMMA" -> flag word. Returns -1 on an unknown name. */
static int64_t parse_flags(const char* str) {
    static const struct { const char* prefix; int shift; } slots[] = {
        { "OUTDEGREES_", 0 }, { "BLOCKS_", 4 }, { "RESIDUALS_", 8 },
        { "REFERENCES_", 12 }, { "BLOCK_COUNT_", 16 }, { "OFFSETS_", 20 } };
    uint32_t flags = 0;
    char tmp[512];
    strncpy(tmp, str, sizeof tmp - 1);
    tmp[sizeof tmp - 1] = 0;
    for (char* tok = strtok(tmp, "|"); tok; tok = strtok(NULL, "|")) {
        while (isspace((unsigned char)*tok)) tok++;
        char* end = tok + strlen(tok);
        while (end > tok && isspace((unsigned char)end[-1])) *--end = 0;
        if (!*tok) continue;
        int found = 0;
        /* BLOCK_COUNT_ must be tried before BLOCKS_ only if prefixes overlapped; they do not. */
        for (unsigned i = 0; i < sizeof slots / sizeof slots[0]; i++) {
            size_t pl = strlen(slots[i].prefix);
            if (!strncmp(tok, slots[i].prefix, pl)) {
                int c = coding_by_name(tok + pl);
                if (c < 0) return -1;
                flags |= (uint32_t)c << slots[i].shift;
                found = 1;
                break;
            }
        }
        if (!found) return -1;
    }
    return flags;
}

/* Minimal java.util.Properties reader: key=value / key:value lines, # and ! comments. */
static int prop_get(const char* text, const char* key, char* out, size_t cap) {
    const char* p = text;
    size_t kl = strlen(key);
    while (*p) {
        const char* eol = strchr(p, '\n');
        if (!eol) eol = p + strlen(p);
        const char* q = p;
        while (q < eol && isspace((unsigned char)*q)) q++;
        if (q < eol && *q != '#' && *q != '!' && (size_t)(eol - q) > kl && !strncmp(q, key, kl)) {
            const char* r = q + kl;
            while (r < eol && (*r == ' ' || *r == '\t')) r++;
            if (r < eol && (*r == '=' || *r == ':')) {
                r++;
                while (r < eol && (*r == ' ' || *r == '\t')) r++;
                size_t len = (size_t)(eol - r);
                while (len && isspace((unsigned char)r[len - 1])) len--;
                if (len >= cap) len = cap - 1;
                memcpy(out, r, len);
                out[len] = 0;
                return 1;
            }
        }
        p = *eol ? eol + 1 : eol;
    }
    return 0;
}

static uint8_t* slurp(const char* path, uint64_t* size, size_t pad) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* b = (uint8_t*)calloc((size_t)sz + pad, 1);
    if (!b) { fclose(f); return NULL; }
    if (sz && fread(b, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(b); return NULL; }
    fclose(f);
    *size = (uint64_t)sz;
    return b;
}

void orc_free(orc_graph* g) {
    if (!g) return;
    free(g->graph);
    free(g->offsets);
    free(g);
}

int orc_from_memory(const uint8_t* graph, uint64_t graph_bytes, const uint64_t* offsets,
                    int32_t n, int64_t m, int32_t window, int32_t maxref, int32_t minlen,
                    int32_t zetak, uint32_t flags, orc_graph** out) {
    orc_graph* g = (orc_graph*)calloc(1, sizeof *g);
    if (!g) return BVGO_ENOMEM;
    g->n = n; g->m = m; g->window = window; g->maxref = maxref; g->minlen = minlen; g->zetak = zetak;
    set_flags(g, flags);
    g->graph = (uint8_t*)calloc(graph_bytes + 16, 1);
    if (!g->graph) { orc_free(g); return BVGO_ENOMEM; }
    memcpy(g->graph, graph, graph_bytes);
    g->graph_bytes = graph_bytes;
    if (offsets) {
        g->offsets = (uint64_t*)malloc(((size_t)n + 1) * sizeof(uint64_t));
        if (!g->offsets) { orc_free(g); return BVGO_ENOMEM; }
        memcpy(g->offsets, offsets, ((size_t)n + 1) * sizeof(uint64_t));
    }
    *out = g;
    return BVGO_OK;
}

int orc_load(const char* basename, int load_offsets, orc_graph** out) {
    char path[4096], val[512];
    uint64_t psz = 0;
    snprintf(path, sizeof path, "%s.properties", basename);
    char* props = (char*)slurp(path, &psz, 1);
    if (!props) return BVGO_EIO;
    orc_graph* g = (orc_graph*)calloc(1, sizeof *g);
    int rc = BVGO_OK;
    /* graphclass check, BVGraph.java:1528 (the "big" package name is accepted) */
    if (!prop_get(props, "graphclass", val, sizeof val) ||
        (strcmp(val, "it.unimi.dsi.webgraph.BVGraph") && strcmp(val, "it.unimi.dsi.big.webgraph.BVGraph"))) { rc = BVGO_EFORMAT; goto fail; }
    if (!prop_get(props, "version", val, sizeof val) || atoi(val) > 0) { rc = BVGO_EFORMAT; goto fail; } /* :1533-1534 */
    val[0] = 0;
    prop_get(props, "compressionflags", val, sizeof val);
    {
        int64_t fl = parse_flags(val);
        if (fl < 0) { rc = BVGO_EFORMAT; goto fail; }
        set_flags(g, (uint32_t)fl);
    }
    if (!prop_get(props, "nodes", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    {
        long long nodes = atoll(val);
        if (nodes > 2147483647LL || nodes < 0) { rc = BVGO_EINVAL; goto fail; } /* :1537 */
        g->n = (int32_t)nodes;
    }
    if (!prop_get(props, "arcs", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    g->m = atoll(val);
    if (!prop_get(props, "windowsize", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    g->window = atoi(val);
    if (!prop_get(props, "maxrefcount", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    g->maxref = atoi(val);
    if (!prop_get(props, "minintervallength", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    g->minlen = atoi(val);
    g->zetak = 3;
    if (prop_get(props, "zetak", val, sizeof val)) g->zetak = atoi(val);

    snprintf(path, sizeof path, "%s.graph", basename);
    g->graph = slurp(path, &g->graph_bytes, 16);
    if (!g->graph) { rc = BVGO_EIO; goto fail; }

    if (load_offsets) {
        /* OffsetsLongIterator, BVGraph.java:907-935: n+1 gaps, running sum. */
        uint64_t osz = 0;
        snprintf(path, sizeof path, "%s.offsets", basename);
        uint8_t* ob = slurp(path, &osz, 16);
        if (!ob) { rc = BVGO_EIO; goto fail; }
        g->offsets = (uint64_t*)malloc(((size_t)g->n + 1) * sizeof(uint64_t));
        ibs_t s = { ob, osz * 8, 0 };
        uint64_t off = 0;
        int err = 0;
        for (int64_t i = 0; i <= g->n; i++) {
            off += read_coded(&s, g->offset_coding, 0, &err); /* readOffset :631-637 (gamma|delta) */
            g->offsets[i] = off;
            if (err || s.pos > s.nbits) { free(ob); rc = err ? err : BVGO_EIO; goto fail; }
        }
        free(ob);
    }
    free(props);
    *out = g;
    return BVGO_OK;
fail:
    free(props);
    orc_free(g);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * The decoder.  decode_record restates BVGraph.successors(x, ibs, window, outd), BVGraph.java:
 * 1032-1133, with the iterator classes flattened:
 *   MaskedIntIterator.java:65-97  -> mask_copy()
 *   IntIntervalSequenceIterator.java:57-95 -> inline expansion
 *   ResidualIntIterator, BVGraph.java:939-991 -> inline loop
 *   MergedIntIterator.java:42-74  -> merge_dedup()
 * ---------------------------------------------------------------------------------------- */
/* Grow-only per-thread scratch so the per-node path does not allocate. */
typedef struct {
    int64_t* block; int64_t block_cap;
    int32_t* buf[4]; int64_t cap[4];   /* copied, extras, intervals, union */
} scratch_t;

static int64_t* scratch_block(scratch_t* sc, int64_t n) {
    if (sc->block_cap < n) { free(sc->block); sc->block = (int64_t*)malloc((size_t)(n + 16) * 2 * sizeof(int64_t)); sc->block_cap = (n + 16) * 2; }
    return sc->block;
}
static int32_t* scratch_buf(scratch_t* sc, int which, int64_t n) {
    if (sc->cap[which] < n + 1) { free(sc->buf[which]); sc->buf[which] = (int32_t*)malloc((size_t)(n + 16) * 2 * sizeof(int32_t)); sc->cap[which] = (n + 16) * 2; }
    return sc->buf[which];
}
static void scratch_free(scratch_t* sc) {
    free(sc->block);
    for (int i = 0; i < 4; i++) free(sc->buf[i]);
    memset(sc, 0, sizeof *sc);
}

typedef struct {
    int32_t** list;   /* W+1 rows */
    int32_t*  cap;
    int32_t*  outd;
    int32_t   size;   /* W+1, BVGraph.java:1041,1142 */
} window_t;

static int64_t decode_random(const orc_graph* g, int32_t x, int32_t** out, int32_t* outcap);

/* MergedIntIterator.nextInt, MergedIntIterator.java:50-74: ascending union, equal heads once. */
static int32_t merge_dedup(const int32_t* a, int32_t na, const int32_t* b, int32_t nb, int32_t* o) {
    int32_t i = 0, j = 0, k = 0;
    while (i < na && j < nb) {
        if (a[i] < b[j]) o[k++] = a[i++];
        else { if (a[i] == b[j]) i++; o[k++] = b[j++]; }
    }
    while (i < na) o[k++] = a[i++];
    while (j < nb) o[k++] = b[j++];
    return k;
}

/* Reads a record whose outdegree d has just been read; writes the d successors to out. */
static int decode_record(const orc_graph* g, int32_t x, ibs_t* s, int32_t d,
                         const window_t* win, scratch_t* sc, int32_t* out) {
    int err = 0;
    if (d == 0) return BVGO_OK; /* :1049 */
    const int32_t W = g->window;
    int32_t ref = -1;
    if (W > 0) { /* :1053, readReference :696-707 */
        uint64_t r = read_coded(s, g->reference_coding, 0, &err);
        if (err) return err;
        if (r > (uint64_t)W) return BVGO_ESTATE; /* :705 */
        ref = (int32_t)r;
    }

    int32_t* copied_list = NULL;
    int32_t copied = 0;
    int rc = BVGO_OK;
    int32_t* extra_list = NULL;

    if (ref > 0) { /* :1058-1071 */
        if (ref > x) return BVGO_EFORMAT; /* would address node x-ref < 0 */
        uint64_t bc64 = read_coded(s, g->block_count_coding, 0, &err);
        if (err) return err;
        if (bc64 > (uint64_t)1 << 31 || s->pos > s->nbits) return BVGO_EIO;
        const int32_t bc = (int32_t)bc64;
        int64_t* block = scratch_block(sc, bc);
        int64_t total = 0, cp = 0;
        for (int32_t i = 0; i < bc; i++) {
            block[i] = (int64_t)read_coded(s, g->block_coding, 0, &err) + (i == 0 ? 0 : 1); /* :1063 */
            total += block[i];
            if ((i & 1) == 0) cp += block[i];
            if (err || s->pos > s->nbits) return err ? err : BVGO_EIO;
        }
        /* the reference list: window row (sequential) or recursion (random access), :1116-1120 */
        const int32_t* parent;
        int32_t dp;
        int32_t* parent_owned = NULL;
        if (win) {
            const int32_t idx = (int32_t)(((int64_t)x - ref + win->size) % win->size); /* :1056 */
            parent = win->list[idx];
            dp = win->outd[idx];
        } else {
            int32_t pcap = 0;
            int64_t r = decode_random(g, x - ref, &parent_owned, &pcap);
            if (r < 0) { free(parent_owned); return (int)r; }
            parent = parent_owned;
            dp = (int32_t)r;
        }
        if ((bc & 1) == 0) cp += dp - total; /* :1069 */
        if (total > dp || cp < 0 || cp > d) { free(parent_owned); return BVGO_EFORMAT; }
        copied = (int32_t)cp;
        copied_list = scratch_buf(sc, 0, copied);
        /* MaskedIntIterator.java:65-97: copy block[0], skip block[1], ...; tail copied iff bc even */
        int32_t k = 0, p = 0;
        for (int32_t i = 0; i < bc; i++) {
            if ((i & 1) == 0) for (int64_t t = 0; t < block[i]; t++) copied_list[k++] = parent[p++];
            else p += (int32_t)block[i];
        }
        if ((bc & 1) == 0) while (p < dp) copied_list[k++] = parent[p++];
        free(parent_owned);
    }

    int64_t extra = (int64_t)d - copied; /* :1070-1072 */
    extra_list = scratch_buf(sc, 1, extra > 0 ? extra : 0);
    int32_t ne = 0;
    int32_t* interval_list = NULL;
    int32_t ni = 0;

    if (extra > 0 && g->minlen != 0) { /* :1076-1096 (always gamma, regardless of flags) */
        const uint64_t ic = ibs_read_gamma(s);
        if (ic > (uint64_t)extra || s->pos > s->nbits) { rc = BVGO_EIO; goto done; }
        if (ic) {
            interval_list = scratch_buf(sc, 2, extra);
            int64_t prev = 0;
            for (uint64_t i = 0; i < ic; i++) {
                int64_t left;
                if (i == 0) left = (int64_t)(int32_t)(nat2int(ibs_read_gamma(s)) + x); /* :1084 (int cast) */
                else left = (int64_t)ibs_read_gamma(s) + prev + 1;                  /* :1091 */
                const int64_t len = (int64_t)ibs_read_gamma(s) + g->minlen;          /* :1085,1092 */
                if (s->pos > s->nbits || len > extra) { rc = BVGO_EIO; goto done; }
                for (int64_t t = 0; t < len; t++) interval_list[ni++] = (int32_t)(left + t);
                prev = left + len;
                extra -= len;
            }
        }
    }

    if (extra > 0) { /* ResidualIntIterator, :939-972 */
        int64_t v = (int64_t)(int32_t)((int64_t)x + nat2int(read_coded(s, g->residual_coding, g->zetak, &err))); /* :954 */
        if (err) { rc = err; goto done; }
        extra_list[ne++] = (int32_t)v;
        for (int64_t i = 1; i < extra; i++) {
            v += (int64_t)read_coded(s, g->residual_coding, g->zetak, &err) + 1; /* :966 */
            extra_list[ne++] = (int32_t)v;
            if (s->pos > s->nbits) { rc = BVGO_EIO; goto done; }
        }
    }
    if (s->pos > s->nbits) { rc = BVGO_EIO; goto done; }

    { /* Merged(Masked, Merged(Intervals, Residuals)), :1103-1126 */
        /* out never aliases a scratch buffer or the parent row, so the last union writes it directly */
        const int32_t* ex = extra_list;
        int32_t nx = ne;
        if (ni) { int32_t* tmp = scratch_buf(sc, 3, (int64_t)ni + ne); nx = merge_dedup(interval_list, ni, extra_list, ne, tmp); ex = tmp; }
        int32_t nf;
        if (copied) nf = merge_dedup(copied_list, copied, ex, nx, out);
        else { memcpy(out, ex, (size_t)nx * sizeof(int32_t)); nf = nx; }
        /* BVGraphNodeIterator.nextInt drains exactly d values, -1 once exhausted (:1210) */
        for (int32_t i = nf; i < d; i++) out[i] = -1;
    }
done:
    return rc;
}

static inline int32_t read_outdegree(const orc_graph* g, ibs_t* s, int* err) { /* :658-664 */
    uint64_t d = read_coded(s, g->outdegree_coding, 0, err);
    if (d > 0x7fffffffULL) { *err = *err ? *err : BVGO_EIO; return 0; }
    return (int32_t)d;
}

int orc_outdegree(const orc_graph* g, int32_t x, int32_t* d) { /* BVGraph.java:857-879 */
    if (x < 0 || x >= g->n) return BVGO_EINVAL;
    if (!g->offsets) return BVGO_ESTATE;
    ibs_t s = { g->graph, g->graph_bytes * 8, g->offsets[x] };
    int err = 0;
    *d = read_outdegree(g, &s, &err);
    if (!err && s.pos > s.nbits) err = BVGO_EIO;
    return err;
}

/* Random access along the reference chain: BVGraph.successors(x) :896-904 with window == null. */
static int64_t decode_random(const orc_graph* g, int32_t x, int32_t** out, int32_t* outcap) {
    if (x < 0 || x >= g->n) return BVGO_EINVAL;   /* :900 */
    if (!g->offsets) return BVGO_EUNSUPPORTED;     /* :901 */
    ibs_t s = { g->graph, g->graph_bytes * 8, g->offsets[x] };
    int err = 0;
    const int32_t d = read_outdegree(g, &s, &err);
    if (err) return err;
    if (s.pos > s.nbits) return BVGO_EIO;
    if (d > *outcap) {
        free(*out);
        *out = (int32_t*)malloc(((size_t)d + 1) * sizeof(int32_t));
        *outcap = d;
    }
    scratch_t sc;
    memset(&sc, 0, sizeof sc);
    int rc = decode_record(g, x, &s, d, NULL, &sc, *out);
    scratch_free(&sc);
    return rc < 0 ? rc : d;
}

int64_t orc_successors(const orc_graph* g, int32_t x, int32_t* out, int64_t cap) {
    int32_t* buf = NULL;
    int32_t bcap = 0;
    int64_t d = decode_random(g, x, &buf, &bcap);
    if (d >= 0) {
        if (d > cap) { free(buf); return BVGO_EINVAL; }
        memcpy(out, buf, (size_t)d * sizeof(int32_t));
    }
    free(buf);
    return d;
}

static void window_free(window_t* w) {
    if (!w->list) return;
    for (int32_t i = 0; i < w->size; i++) free(w->list[i]);
    free(w->list); free(w->cap); free(w->outd);
}

static void window_reserve(window_t* w, int32_t idx, int32_t d) {
    if (w->cap[idx] < d) {
        free(w->list[idx]);
        w->list[idx] = (int32_t*)malloc(((size_t)d + 1) * sizeof(int32_t));
        w->cap[idx] = d;
    }
}

/* Shared sequential driver (BVGraphNodeIterator ctor :1164-1186 + nextInt :1200-1213). mode:
 * 0 = materialise into out/out_off, 1 = consume-only checksum, 2 = record offsets. */
static int64_t sequential(const orc_graph* g, int32_t from, int32_t to, int mode,
                          int64_t* out_off, int32_t* out, int64_t cap, uint64_t* checksum, uint64_t* offs) {
    if (from < 0 || from > g->n || to < from || to > g->n) return BVGO_EINVAL; /* :1165 */
    if (from != 0 && !g->offsets) return BVGO_ESTATE;                          /* :1174 */
    window_t w;
    w.size = g->window + 1;
    w.list = (int32_t**)calloc((size_t)w.size, sizeof(int32_t*));
    w.cap = (int32_t*)calloc((size_t)w.size, sizeof(int32_t));
    w.outd = (int32_t*)calloc((size_t)w.size, sizeof(int32_t));
    int64_t rc = 0;
    scratch_t sc;
    memset(&sc, 0, sizeof sc);
    /* seed the window by random access, :1173-1183 */
    if (from != 0) {
        for (int32_t i = 1; i < (from + 1 < w.size ? from + 1 : w.size); i++) {
            const int32_t idx = (int32_t)(((int64_t)from - i + w.size) % w.size);
            int64_t d = decode_random(g, from - i, &w.list[idx], &w.cap[idx]);
            if (d < 0) { rc = d; goto done; }
            w.outd[idx] = (int32_t)d;
        }
    }
    {
        ibs_t s = { g->graph, g->graph_bytes * 8, from ? g->offsets[from] : 0 };
        int64_t arcs = 0;
        uint64_t cs = 0;
        if (out_off) out_off[0] = 0;
        for (int32_t x = from; x < to; x++) {
            const int32_t idx = x % w.size; /* :1204 */
            int err = 0;
            if (offs) offs[x - from] = s.pos;
            const int32_t d = read_outdegree(g, &s, &err); /* :1048 */
            if (err) { rc = err; goto done; }
            if (s.pos > s.nbits) { rc = BVGO_EIO; goto done; }
            window_reserve(&w, idx, d);
            int r = decode_record(g, x, &s, d, &w, &sc, w.list[idx]);
            if (r < 0) { rc = r; goto done; }
            w.outd[idx] = d;
            if (mode == 0) {
                if (out) {
                    if (arcs + d > cap) { rc = BVGO_EINVAL; goto done; }
                    memcpy(out + arcs, w.list[idx], (size_t)d * sizeof(int32_t));
                }
                out_off[x - from + 1] = arcs + d;
            } else if (mode == 1) {
                const uint64_t base = (uint64_t)(uint32_t)x * 0x9E3779B97F4A7C15ULL;
                for (int32_t j = 0; j < d; j++) cs ^= base + (uint64_t)(uint32_t)w.list[idx][j];
            }
            arcs += d;
        }
        if (offs) offs[to - from] = s.pos;
        if (checksum) *checksum = cs;
        rc = arcs;
    }
done:
    window_free(&w);
    scratch_free(&sc);
    return rc;
}

int64_t orc_decode_range(const orc_graph* g, int32_t from, int32_t to,
                         int64_t* out_off, int32_t* out, int64_t cap) {
    return sequential(g, from, to, 0, out_off, out, cap, NULL, NULL);
}

int orc_scan_range(const orc_graph* g, int32_t from, int32_t to, int64_t* arcs, uint64_t* checksum) {
    int64_t r = sequential(g, from, to, 1, NULL, NULL, 0, checksum, NULL);
    if (r < 0) return (int)r;
    *arcs = r;
    return BVGO_OK;
}

int orc_rebuild_offsets(const orc_graph* g, uint64_t* out) {
    int64_t r = sequential(g, 0, g->n, 2, NULL, NULL, 0, NULL, out);
    return r < 0 ? (int)r : BVGO_OK;
}

/* Header-only walk up the chain of x: record bits of x and of each ancestor. */
int64_t orc_chain_bits(const orc_graph* g, int32_t x, int32_t* depth) {
    if (x < 0 || x >= g->n) return BVGO_EINVAL;
    if (!g->offsets) return BVGO_EUNSUPPORTED;
    int64_t bits = 0;
    int32_t dep = 0;
    for (;;) {
        bits += (int64_t)(g->offsets[x + 1] - g->offsets[x]);
        ibs_t s = { g->graph, g->graph_bytes * 8, g->offsets[x] };
        int err = 0;
        const int32_t d = read_outdegree(g, &s, &err);
        if (err) return err;
        if (d == 0 || g->window == 0) break;
        const uint64_t ref = read_coded(&s, g->reference_coding, 0, &err);
        if (err) return err;
        if (ref == 0) break;
        if (ref > (uint64_t)g->window) return BVGO_ESTATE;
        if (ref > (uint64_t)x) return BVGO_EFORMAT;
        x -= (int32_t)ref;
        dep++;
    }
    if (depth) *depth = dep;
    return bits;
}

/* Root of x's reference chain: the node the recursion of successors(x) bottoms out at (BVGraph.java:1110-1121). */
int32_t orc_chain_root(const orc_graph* g, int32_t x) {
    if (x < 0 || x >= g->n) return BVGO_EINVAL;
    if (!g->offsets) return BVGO_EUNSUPPORTED;
    for (;;) {
        ibs_t s = { g->graph, g->graph_bytes * 8, g->offsets[x] };
        int err = 0;
        const int32_t d = read_outdegree(g, &s, &err);
        if (err) return err;
        if (d == 0 || g->window == 0) return x;
        const uint64_t ref = read_coded(&s, g->reference_coding, 0, &err);
        if (err) return err;
        if (ref == 0) return x;
        if (ref > (uint64_t)g->window) return BVGO_ESTATE;
        if (ref > (uint64_t)x) return BVGO_EFORMAT;
        x -= (int32_t)ref;
    }
}

/* ------------------------------------------------------------------------------------------
 * Arc labels: labelling/BitStreamArcLabelledImmutableGraph.java.  PARITY UNPINNED: the reference
 * ships no .labels fixture (its test writes them with dsiutils' OutputBitStream at run time,
 * test/.../labelling/BitStreamArcLabelledGraphTest.java:131-203) and no JVM exists in this image;
 * the restatement is anchored on that test's layout and on the three fromBitStream methods.
 * ---------------------------------------------------------------------------------------- */

/* labelspec = <class>(<key>[,<width>]) (ObjectParser.fromSpec call, :409-425). */
static int parse_labelspec(const char* spec, int* kind, int* width) {
    char cls[256];
    const char* par = strchr(spec, '(');
    size_t cl = par ? (size_t)(par - spec) : strlen(spec);
    while (cl && isspace((unsigned char)spec[cl - 1])) cl--;
    size_t start = cl;
    while (start && spec[start - 1] != '.') start--;
    if (cl - start >= sizeof cls) return BVGO_EFORMAT;
    memcpy(cls, spec + start, cl - start);
    cls[cl - start] = 0;
    *width = 0;
    if (!strcmp(cls, "GammaCodedIntLabel")) { *kind = ORC_LABEL_GAMMA; return par ? BVGO_OK : BVGO_EFORMAT; }
    if (!strcmp(cls, "FixedWidthIntLabel")) *kind = ORC_LABEL_FIXED;
    else if (!strcmp(cls, "FixedWidthIntListLabel")) *kind = ORC_LABEL_FIXED_LIST;
    else return BVGO_EUNSUPPORTED;
    const char* comma = par ? strchr(par, ',') : NULL;
    if (!comma) return BVGO_EFORMAT;
    char* end = NULL;
    const long w = strtol(comma + 1, &end, 10);
    if (end == comma + 1) return BVGO_EFORMAT;
    if (w < 0 || w > 31) return BVGO_EINVAL; /* FixedWidthIntLabel.java:41, FixedWidthIntListLabel.java:44 */
    *width = (int)w;
    return BVGO_OK;
}

void orc_labels_free(orc_labels* l) {
    if (!l) return;
    free(l->labels);
    free(l->offsets);
    free(l);
}

/* load(), :385-470, for a graph of n nodes: labelspec -> prototype, .labels bytes, .labeloffsets = n+1 gamma gaps summed
 * (LabelOffsetsLongIterator, :330-358). */
int orc_labels_load(const char* basename, int32_t n, orc_labels** out) {
    char path[4096], val[1024];
    uint64_t psz = 0;
    snprintf(path, sizeof path, "%s.properties", basename);
    char* props = (char*)slurp(path, &psz, 1);
    if (!props) return BVGO_EIO;
    orc_labels* l = (orc_labels*)calloc(1, sizeof *l);
    int rc = BVGO_OK;
    if (!prop_get(props, "labelspec", val, sizeof val)) { rc = BVGO_EIO; goto fail; } /* :409 */
    rc = parse_labelspec(val, &l->kind, &l->width);
    if (rc) goto fail;
    l->n = n;
    snprintf(path, sizeof path, "%s.labels", basename);
    l->labels = slurp(path, &l->label_bytes, 16);
    if (!l->labels) { rc = BVGO_EIO; goto fail; }
    {
        uint64_t osz = 0;
        snprintf(path, sizeof path, "%s.labeloffsets", basename);
        uint8_t* ob = slurp(path, &osz, 16);
        if (!ob) { rc = BVGO_EIO; goto fail; }
        l->offsets = (uint64_t*)malloc(((size_t)n + 1) * sizeof(uint64_t));
        ibs_t s = { ob, osz * 8, 0 };
        uint64_t off = 0;
        for (int64_t i = 0; i <= n; i++) {
            off += ibs_read_gamma(&s);
            l->offsets[i] = off;
            if (s.pos > s.nbits) { free(ob); rc = BVGO_EIO; goto fail; }
        }
        free(ob);
        if (l->offsets[n] > l->label_bytes * 8) { rc = BVGO_EIO; goto fail; }
    }
    free(props);
    *out = l;
    return BVGO_OK;
fail:
    free(props);
    orc_labels_free(l);
    return rc;
}

/* BitStreamLabelledArcIterator (:225-262): position(offset(x)), then one fromBitStream per successor of x
 * (GammaCodedIntLabel.java:52-56 readGamma; FixedWidthIntLabel.java:69-73 readInt(width);
 * FixedWidthIntListLabel.java:72-78 readGamma then length x readInt(width)).  d = outdegree of x in the
 * underlying graph.  list_off (d+1 entries, may be NULL) = where each arc's values start; returns the
 * number of values (== d for the integer labels) or an error; values may be NULL (count only). */
int64_t orc_labels_node(const orc_labels* l, int32_t x, int32_t d, int64_t* list_off, int32_t* values, int64_t cap) {
    if (x < 0 || x >= l->n || d < 0) return BVGO_EINVAL;
    ibs_t s = { l->labels, l->label_bytes * 8, l->offsets[x] };
    int64_t nv = 0;
    for (int32_t k = 0; k < d; k++) {
        if (list_off) list_off[k] = nv;
        uint64_t len = 1;
        if (l->kind == ORC_LABEL_FIXED_LIST) len = ibs_read_gamma(&s);
        if (len > 0x7fffffffULL || s.pos > s.nbits) return BVGO_EIO;
        for (uint64_t i = 0; i < len; i++) {
            const uint64_t v = l->kind == ORC_LABEL_GAMMA ? ibs_read_gamma(&s) : ibs_read_bits(&s, l->width);
            if (s.pos > s.nbits) return BVGO_EIO;
            if (values) {
                if (nv >= cap) return BVGO_ENOMEM;
                values[nv] = (int32_t)v;
            }
            nv++;
        }
    }
    if (list_off) list_off[d] = nv;
    return nv;
}

/* Sequential pass over the labels of nodes [from, to) (what ArcLabelledNodeIterator does: the label stream read front
 * to back, one fromBitStream per arc), integer labels only; row_off = CSR row offsets of the underlying graph.
 * values may be NULL (consume only); returns the number of labels or an error; *sum gets the sum of the values. */
int64_t orc_labels_range(const orc_labels* l, int32_t from, int32_t to, const int64_t* row_off, int32_t* values, int64_t cap, uint64_t* sum) {
    if (from < 0 || to < from || to > l->n || l->kind == ORC_LABEL_FIXED_LIST) return BVGO_EINVAL;
    ibs_t s = { l->labels, l->label_bytes * 8, l->offsets[from] };
    const int64_t arcs = row_off[to] - row_off[from];
    if (values && cap < arcs) return BVGO_ENOMEM;
    uint64_t acc = 0;
    for (int64_t j = 0; j < arcs; j++) {
        const uint64_t v = l->kind == ORC_LABEL_GAMMA ? ibs_read_gamma(&s) : ibs_read_bits(&s, l->width);
        if (values) values[j] = (int32_t)v;
        acc += v;
    }
    if (s.pos > s.nbits) return BVGO_EIO;
    if (sum) *sum = acc;
    return arcs;
}

/* ------------------------------------------------------------------------------------------
 * EFGraph (EFGraph.java): the quasi-succinct format.  PARITY UNPINNED: the reference ships no
 * EFGraph fixture (EFGraphTest.java stores its graphs at run time) and no JVM exists here; the
 * restatement follows LongWordBitReader (:892-1036) and EliasFanoSuccessorReader (:1064-1145).
 * The stream is LSB-first in 64-bit words: bit i is bit i % 64 of word i / 64.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t ef_bits(const uint64_t* w, uint64_t pos, int width) { /* extract(position), :981-1000 */
    if (width == 0) return 0;
    const uint64_t i = pos >> 6;
    const int s = (int)(pos & 63);
    uint64_t v = w[i] >> s;
    if (s && s + width > 64) v |= w[i + 1] << (64 - s);
    return width == 64 ? v : v & ((1ULL << width) - 1);
}
static inline int ef_msb(uint64_t x) { return x ? 63 - __builtin_clzll(x) : -1; }
static inline int ef_lower_bits(uint64_t length, uint64_t ub) { /* :145-147 */
    if (length == 0) return 0;
    const int m = ef_msb(ub / length);
    return m < 0 ? 0 : m;
}
static inline int ef_ceil_log2(uint64_t x) { return x <= 1 ? 0 : 64 - __builtin_clzll(x - 1); }

void orc_ef_free(orc_efgraph* g) {
    if (!g) return;
    free(g->words);
    free(g->offsets);
    free(g);
}

int orc_ef_load(const char* basename, orc_efgraph** out) { /* loadInternal, :709-790 */
    char path[4096], val[512];
    uint64_t psz = 0;
    snprintf(path, sizeof path, "%s.properties", basename);
    char* props = (char*)slurp(path, &psz, 1);
    if (!props) return BVGO_EIO;
    orc_efgraph* g = (orc_efgraph*)calloc(1, sizeof *g);
    int rc = BVGO_OK, big = 0;
    if (!prop_get(props, "graphclass", val, sizeof val) ||
        (strcmp(val, "it.unimi.dsi.webgraph.EFGraph") && strcmp(val, "it.unimi.dsi.big.webgraph.EFGraph"))) { rc = BVGO_EIO; goto fail; }
    if (!prop_get(props, "version", val, sizeof val) || atoi(val) > 0) { rc = BVGO_EIO; goto fail; }
    if (!prop_get(props, "nodes", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    if (atoll(val) > 2147483647LL || atoll(val) < 0) { rc = BVGO_EINVAL; goto fail; }
    g->n = (int32_t)atoll(val);
    if (!prop_get(props, "arcs", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    g->m = atoll(val);
    g->upper_bound = g->n;
    if (prop_get(props, "upperbound", val, sizeof val)) g->upper_bound = atoi(val);
    if (!prop_get(props, "quantum", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    {
        const long long quantum = atoll(val);
        g->log2_quantum = ef_msb((uint64_t)quantum);
        if (quantum <= 0 || (1LL << g->log2_quantum) != quantum) { rc = BVGO_EINVAL; goto fail; } /* "Illegal quantum" */
    }
    if (!prop_get(props, "byteorder", val, sizeof val)) { rc = BVGO_EFORMAT; goto fail; }
    if (!strcmp(val, "BIG_ENDIAN")) big = 1;
    else if (strcmp(val, "LITTLE_ENDIAN")) { rc = BVGO_EINVAL; goto fail; } /* "Unknown byte order" */
    {
        uint64_t bytes = 0;
        snprintf(path, sizeof path, "%s.graph", basename);
        uint8_t* b = slurp(path, &bytes, 32);
        if (!b) { rc = BVGO_EIO; goto fail; }
        g->nwords = bytes / 8;
        g->words = (uint64_t*)b; /* calloc'ed: aligned; two padding words follow */
        if (big) for (uint64_t i = 0; i < g->nwords; i++) g->words[i] = __builtin_bswap64(g->words[i]);
    }
    {
        uint64_t osz = 0;
        snprintf(path, sizeof path, "%s.offsets", basename);
        uint8_t* ob = slurp(path, &osz, 16);
        if (!ob) { rc = BVGO_EIO; goto fail; }
        g->offsets = (uint64_t*)malloc(((size_t)g->n + 1) * sizeof(uint64_t));
        ibs_t s = { ob, osz * 8, 0 };
        uint64_t off = 0;
        for (int64_t i = 0; i <= g->n; i++) { /* OffsetsLongIterator, :641-672: delta-coded gaps */
            off += ibs_read_delta(&s);
            g->offsets[i] = off;
            if (s.pos > s.nbits) { free(ob); rc = BVGO_EIO; goto fail; }
        }
        free(ob);
        if (g->offsets[g->n] > g->nwords * 64) { rc = BVGO_EIO; goto fail; }
    }
    free(props);
    *out = g;
    return BVGO_OK;
fail:
    free(props);
    orc_ef_free(g);
    return rc;
}

/* readGamma at a bit position (:1002-1036): trailing zeros = msb, then msb bits. */
static uint64_t ef_read_gamma(const uint64_t* w, uint64_t* pos) {
    uint64_t p = *pos;
    int msb = 0;
    for (;;) {
        const uint64_t v = w[p >> 6] >> (p & 63);
        if (v) { const int z = __builtin_ctzll(v); msb += z; p += (uint64_t)z + 1; break; }
        msb += 64 - (int)(p & 63);
        p = ((p >> 6) + 1) << 6;
        if (msb > 64) { *pos = p; return ~0ULL; }
    }
    const uint64_t low = ef_bits(w, p, msb);
    *pos = p + (uint64_t)msb;
    return (low | (1ULL << msb)) - 1;
}

int orc_ef_outdegree(const orc_efgraph* g, int32_t x, int32_t* d) { /* :1054-1060 */
    if (x < 0 || x >= g->n) return BVGO_EINVAL;
    uint64_t pos = g->offsets[x];
    const uint64_t v = ef_read_gamma(g->words, &pos);
    if (v > 0x7fffffffULL) return BVGO_EIO;
    *d = (int32_t)v;
    return BVGO_OK;
}

/* successors(x) drained (EliasFanoSuccessorReader, :1100-1145): returns d or an error. */
int64_t orc_ef_successors(const orc_efgraph* g, int32_t x, int32_t* out, int64_t cap) {
    if (x < 0 || x >= g->n) return BVGO_EINVAL;
    uint64_t pos = g->offsets[x];
    const uint64_t d = ef_read_gamma(g->words, &pos);
    if (d > 0x7fffffffULL) return BVGO_EIO;
    if (out && (int64_t)d > cap) return BVGO_ENOMEM;
    if (d == 0 || !out) return (int64_t)d;
    const uint64_t ub = (uint64_t)g->upper_bound, len = d + 1;
    const int l = ef_lower_bits(len, ub);
    const int psize = ef_ceil_log2(len + (ub >> l));
    const uint64_t npointers = (ub >> l) >> g->log2_quantum;
    const uint64_t lower_start = pos + (uint64_t)psize * npointers, upper_start = lower_start + (uint64_t)l * len;
    uint64_t curr = upper_start >> 6;
    uint64_t window = g->words[curr] & (~0ULL << (upper_start & 63));
    for (uint64_t k = 0; k < d; k++) {
        while (window == 0) {
            if (++curr >= g->nwords + 2) return BVGO_EIO;
            window = g->words[curr];
        }
        const uint64_t upper = curr * 64 + (uint64_t)__builtin_ctzll(window) - k - upper_start;
        window &= window - 1;
        out[k] = (int32_t)((upper << l) | ef_bits(g->words, lower_start + (uint64_t)l * k, l));
    }
    return (int64_t)d;
}
