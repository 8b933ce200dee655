/*
 * gbin_oracle.c — TEST INFRASTRUCTURE ONLY (see gbin_oracle.h).
 *
 * Plain-C restatement of the reference hot path.  Written from the behaviour of
 * /root/reference/binning.c (file:line cited per function); flat arrays and integer codes instead
 * of the reference's string-keyed chained hash tables and linked lists.
 * Parity: pinned against the reference binary (oracle/_ref) — see tests/test_oracle_pins.py.
 */
#include "gbin_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- scalar helpers */

/* binning.c:91-111 — T0 G1 C2 A3, anything else 3 */
int orc_getval(char c)
{
    if (c == 'T') return 0;
    if (c == 'G') return 1;
    if (c == 'C') return 2;
    return 3;
}

/* binning.c:69-88 — inverse map, out of range -> 'A' */
char orc_getbp(int bp)
{
    static const char tab[4] = {'T', 'G', 'C', 'A'};
    return (bp >= 0 && bp <= 3) ? tab[bp] : 'A';
}

/* binning.c:114-124 — base-4 positional score in int arithmetic */
int orc_getscore(const char *s)
{
    int score = 0;
    for (; *s != '\0'; s++) score = score * 4 + orc_getval(*s);
    return score;
}

/* ---------------------------------------------------------------- process_read */

static void kmer_code(const char *p, int K, uint64_t *hi, uint64_t *lo)
{
    uint64_t h = 0, l = 0;
    for (int t = 0; t < K; t++) {
        h = (h << 2) | (l >> 62);
        l = (l << 2) | (uint64_t)orc_getval(p[t]);
    }
    *hi = h;
    *lo = l;
}

/* binning.c:902-1073.  Pointers become positions: `kmer` = i, `signature` = sig (NULL = -1).
 * Both branches of the reference are restated, including the `else` branch whose inner loop
 * (binning.c:997) runs zero times whenever K >= 2M. */
size_t orc_process_read(const char *read, int len, int K, int M, uint32_t arrival,
                        orc_tuple *tuples, orc_window *windows)
{
    char mmer[32];
    int score = 0, rev_score = 0, max_score = 0, msb = 0;
    int is_rev = 0;
    long sig = -1;
    int top = 1; /* power_val[MMER_SIZE-1], binning.c:17 */
    for (int t = 1; t < M; t++) top *= 4;
    const int full = top * 4 - 1;
    size_t n = 0;

    for (int i = 0; i < len - K + 1; i++) {
        const char *kmer = read + i;
        if ((long)i > sig) { /* binning.c:922: kmer > signature */
            score = rev_score = max_score = 0;
            for (int j = 0; j < M; j++) { /* :931-936 */
                mmer[j] = kmer[j];
                score = score * 4 + orc_getval(kmer[j]);
                rev_score = rev_score * 4 + 3 - orc_getval(kmer[j]);
            }
            if (score > rev_score) { max_score = score; is_rev = 0; } /* :940-949 */
            else { max_score = rev_score; is_rev = 1; }
            sig = i;
            msb = 0;
            int j = M;
            while (j < K) { /* :955-988 */
                score = (score - orc_getval(mmer[msb]) * top) * 4 + orc_getval(kmer[j]);
                rev_score = (rev_score - (3 - orc_getval(mmer[msb])) * top) * 4 + 3 - orc_getval(kmer[j]);
                mmer[msb] = kmer[j];
                msb = (msb + 1) % M;
                j++;
                int best = score > rev_score ? score : rev_score;
                if (best > max_score) { /* strict: leftmost maximum wins */
                    if (score > rev_score) { max_score = score; is_rev = 0; }
                    else { max_score = rev_score; is_rev = 1; }
                    sig = i + j - M;
                }
            }
        } else { /* :992-1021 */
            for (int j = K - M; j < M; j++) {
                mmer[j] = kmer[j];
                score = score * 4 + orc_getval(kmer[j]);
                rev_score = rev_score * 4 + 3 - orc_getval(kmer[j]);
            }
            int best = score > rev_score ? score : rev_score;
            if (best > max_score) {
                if (score > rev_score) { max_score = score; is_rev = 0; }
                else { max_score = rev_score; is_rev = 1; }
                sig = i + K - M;
            }
        }
        /* :1023-1040 — keys; in code space the complement of an n-base code c is (4^n - 1) - c */
        if (tuples) {
            uint64_t hi, lo;
            kmer_code(kmer, K, &hi, &lo);
            if (is_rev) {
                hi = ~hi;
                lo = ~lo;
                if (K < 32) { lo &= (((uint64_t)1 << (2 * K)) - 1); hi = 0; }
                else if (K == 32) { hi = 0; }
                else if (K < 64) { hi &= (((uint64_t)1 << (2 * (K - 32))) - 1); }
            }
            /* stored m-mer key: signature chars, complemented when is_rev */
            uint32_t mc = 0;
            for (int t = 0; t < M; t++) {
                int v = orc_getval(read[sig + t]);
                mc = mc * 4 + (uint32_t)(is_rev ? 3 - v : v);
            }
            tuples[n].mmer = mc;
            tuples[n].arrival = arrival;
            tuples[n].khi = hi;
            tuples[n].klo = lo;
        }
        if (windows) {
            windows[n].sig_pos = (int32_t)sig;
            windows[n].is_rev = is_rev;
            windows[n].mmer = (uint32_t)max_score;
        }
        (void)full;
        n++;
    }
    return n;
}

/* ---------------------------------------------------------------- main's fgets loop */

/* binning.c:1154-1166.  fgets(buf, R) stores at most R-1 bytes, stopping after a '\n'.
 * Then `read[--len] = '\0'` drops the last byte whatever it is, and read_id++ happens for every
 * fgets return (so an over-long line is split into several "reads", and the lone "\n" left over
 * after an exactly-(R-1)-byte line becomes an empty read that still owns an id). */
size_t orc_fgets_split(const char *data, size_t size, int R, uint64_t **starts_out, uint32_t **lens_out)
{
    size_t cap = 1024, n = 0;
    uint64_t *starts = malloc(cap * sizeof *starts);
    uint32_t *lens = malloc(cap * sizeof *lens);
    size_t pos = 0;
    while (pos < size) {
        size_t take = 0;
        while (take < (size_t)(R - 1) && pos + take < size) {
            char c = data[pos + take];
            take++;
            if (c == '\n') break;
        }
        /* NUL bytes inside the file would shorten strlen(); the fixtures have none. */
        if (n == cap) {
            cap *= 2;
            starts = realloc(starts, cap * sizeof *starts);
            lens = realloc(lens, cap * sizeof *lens);
        }
        starts[n] = pos;
        lens[n] = (uint32_t)(take - 1);
        n++;
        pos += take;
    }
    *starts_out = starts;
    *lens_out = lens;
    return n;
}

/* ---------------------------------------------------------------- table build + prune */

static int tuple_cmp(const void *a, const void *b)
{
    const orc_tuple *x = a, *y = b;
    if (x->mmer != y->mmer) return x->mmer < y->mmer ? -1 : 1;
    if (x->khi != y->khi) return x->khi < y->khi ? -1 : 1;
    if (x->klo != y->klo) return x->klo < y->klo ? -1 : 1;
    if (x->arrival != y->arrival) return x->arrival < y->arrival ? -1 : 1;
    return 0;
}

static int same_key(const orc_tuple *a, const orc_tuple *b)
{
    return a->mmer == b->mmer && a->khi == b->khi && a->klo == b->klo;
}

size_t orc_scan_all(const char *data, const uint64_t *starts, const uint32_t *lens, size_t n_reads,
                    int K, int M, orc_tuple *tuples, orc_window *windows)
{
    size_t n = 0;
    for (size_t r = 0; r < n_reads; r++) {
        n += orc_process_read(data + starts[r], (int)lens[r], K, M, (uint32_t)r,
                              tuples ? tuples + n : NULL, windows ? windows + n : NULL);
    }
    return n;
}

/* The grouping the two-level zhash performs implicitly (binning.c:1044-1069: one ll_node per
 * instance, newest id at the head) and prune_data / prune_kmers (binning.c:1085-1144: keep iff list
 * length > ABUNDANCE_CUTOFF; drop emptied buckets), over tuples given in arrival order.
 * Sorts `t` in place. ids maps arrival -> read id (NULL: id = id_base + arrival). */
int orc_group_tuples(orc_tuple *t, size_t n, const int32_t *ids, int32_t id_base, int K, int M, int cutoff, orc_result *out)
{
    memset(out, 0, sizeof *out);
    out->K = K;
    out->M = M;
    out->cutoff = cutoff;
    out->kw = K <= 32 ? 1 : 2;
    qsort(t, n, sizeof *t, tuple_cmp);

    /* pass 1: count */
    uint64_t distinct = 0, S = 0, NS = 0, B = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && same_key(&t[i], &t[j])) j++;
        distinct++;
        if (cutoff < 0 || (long long)(j - i) > cutoff) {
            S++;
            NS += j - i;
        }
        i = j;
    }
    out->n_instances = n;
    out->n_distinct = distinct;
    out->n_kmers = S;
    out->n_ids = NS;
    out->kmer_codes = malloc((S * out->kw + 1) * sizeof(uint64_t));
    out->kmer_id_off = malloc((S + 1) * sizeof(uint64_t));
    out->read_ids = malloc((NS + 1) * sizeof(int32_t));
    out->mmer_codes = malloc((S + 1) * sizeof(uint32_t));
    out->mmer_kmer_off = malloc((S + 2) * sizeof(uint64_t));
    /* pass 2: emit */
    uint64_t s = 0, idp = 0;
    int have_prev = 0;
    uint32_t prev_mmer = 0;
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && same_key(&t[i], &t[j])) j++;
        if (cutoff < 0 || (long long)(j - i) > cutoff) {
            if (!have_prev || prev_mmer != t[i].mmer) {
                out->mmer_codes[B] = t[i].mmer;
                out->mmer_kmer_off[B] = s;
                B++;
                have_prev = 1;
                prev_mmer = t[i].mmer;
            }
            if (out->kw == 1) out->kmer_codes[s] = t[i].klo;
            else { out->kmer_codes[2 * s] = t[i].khi; out->kmer_codes[2 * s + 1] = t[i].klo; }
            out->kmer_id_off[s] = idp;
            for (size_t q = j; q-- > i;) { /* newest (latest arrival) first */
                uint32_t a = t[q].arrival;
                out->read_ids[idp++] = ids ? ids[a] : id_base + (int32_t)a;
            }
            s++;
        }
        i = j;
    }
    out->kmer_id_off[s] = idp;
    out->mmer_kmer_off[B] = s;
    out->n_buckets = B;
    return 0;
}

/* process_read for every read, then grouping + prune. */
int orc_run(const char *data, const uint64_t *starts, const uint32_t *lens, size_t n_reads,
            const int32_t *ids, int K, int M, int cutoff, orc_result *out)
{
    size_t total = 0;
    for (size_t r = 0; r < n_reads; r++)
        if ((int)lens[r] >= K) total += lens[r] - K + 1;
    orc_tuple *t = malloc((total ? total : 1) * sizeof *t);
    if (!t) return -1;
    size_t n = orc_scan_all(data, starts, lens, n_reads, K, M, t, NULL);
    if (n != total) { free(t); return -2; }
    int rc = orc_group_tuples(t, n, ids, 0, K, M, cutoff, out);
    free(t);
    return rc;
}

void orc_result_free(orc_result *r)
{
    free(r->mmer_codes);
    free(r->mmer_kmer_off);
    free(r->kmer_codes);
    free(r->kmer_id_off);
    free(r->read_ids);
    memset(r, 0, sizeof *r);
}

/* ---------------------------------------------------------------- dumps */

void orc_decode(uint64_t hi, uint64_t lo, int n, char *dst)
{
    for (int t = n - 1; t >= 0; t--) {
        dst[t] = orc_getbp((int)(lo & 3));
        lo = (lo >> 2) | (hi << 62);
        hi >>= 2;
    }
    dst[n] = '\0';
}

int orc_dump(const orc_result *r, FILE *f)
{
    char mm[40], km[80];
    for (uint64_t b = 0; b < r->n_buckets; b++) {
        orc_decode(0, r->mmer_codes[b], r->M, mm);
        for (uint64_t s = r->mmer_kmer_off[b]; s < r->mmer_kmer_off[b + 1]; s++) {
            if (r->kw == 1) orc_decode(0, r->kmer_codes[s], r->K, km);
            else orc_decode(r->kmer_codes[2 * s], r->kmer_codes[2 * s + 1], r->K, km);
            fputs(mm, f);
            fputc(' ', f);
            fputs(km, f);
            for (uint64_t q = r->kmer_id_off[s]; q < r->kmer_id_off[s + 1]; q++) fprintf(f, " %d", r->read_ids[q]);
            fputc('\n', f);
        }
    }
    return 0;
}

/* String-faithful path: literal key strings as binning.c:1023-1040 builds them. */
typedef struct str_tuple {
    char key[112]; /* "<mmer> <kmer>" */
    uint32_t arrival;
} str_tuple;

static int str_cmp(const void *a, const void *b)
{
    const str_tuple *x = a, *y = b;
    int c = strcmp(x->key, y->key);
    if (c) return c;
    return x->arrival < y->arrival ? -1 : (x->arrival > y->arrival);
}

int orc_dump_strings(const char *data, const uint64_t *starts, const uint32_t *lens, size_t n_reads,
                     const int32_t *ids, int K, int M, int cutoff, FILE *f)
{
    size_t total = 0;
    for (size_t r = 0; r < n_reads; r++)
        if ((int)lens[r] >= K) total += lens[r] - K + 1;
    str_tuple *t = malloc((total ? total : 1) * sizeof *t);
    orc_window *w = malloc((total ? total : 1) * sizeof *w);
    if (!t || !w) return -1;
    size_t n = 0;
    for (size_t r = 0; r < n_reads; r++) {
        const char *read = data + starts[r];
        size_t nw = orc_process_read(read, (int)lens[r], K, M, (uint32_t)r, NULL, w);
        for (size_t i = 0; i < nw; i++, n++) {
            char *k = t[n].key;
            const char *sp = read + w[i].sig_pos;
            for (int q = 0; q < M; q++) k[q] = w[i].is_rev ? orc_getbp(3 - orc_getval(sp[q])) : sp[q];
            k[M] = ' ';
            for (int q = 0; q < K; q++) k[M + 1 + q] = w[i].is_rev ? orc_getbp(3 - orc_getval(read[i + q])) : read[i + q];
            k[M + 1 + K] = '\0';
            t[n].arrival = (uint32_t)r;
        }
    }
    qsort(t, n, sizeof *t, str_cmp);
    for (size_t i = 0; i < n;) {
        size_t j = i + 1;
        while (j < n && strcmp(t[i].key, t[j].key) == 0) j++;
        if (cutoff < 0 || (long long)(j - i) > cutoff) {
            fputs(t[i].key, f);
            for (size_t q = j; q-- > i;) fprintf(f, " %d", ids ? ids[t[q].arrival] : (int32_t)t[q].arrival);
            fputc('\n', f);
        }
        i = j;
    }
    free(t);
    free(w);
    return 0;
}
