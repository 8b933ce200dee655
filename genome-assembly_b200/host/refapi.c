/*
 * refapi.c — plain-C host side of libgbin.so: the reference's own entry points for the binning path
 * (process_read, prune_data, getval, getbp, getscore — binning.c:902,1130,91,69,114) implemented on
 * top of the gbin_* C ABI (CUDA behind it), plus the host-only helpers: main's fgets reader,
 * table dumps and the adapter that materialises the reference's ZHashTable / ll_node pointer graph.
 *
 * Nothing here computes the table on the CPU: process_read only stages the read, prune_data runs the
 * GPU pipeline (gbin_bin_reads_host) and then copies the flat result into the pointer graph.
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/gbin.h"
#include "../../include/gbin_ref_types.h"

/* ------------------------------------------------------------------ scalar helpers (host, pure) */

/* binning.c:91-111 */
int getval(char c)
{
    switch (c) {
    case 'T': return 0;
    case 'G': return 1;
    case 'C': return 2;
    default: return 3; /* 'A' and everything else */
    }
}

/* binning.c:69-88 */
char getbp(int bp)
{
    switch (bp) {
    case 0: return 'T';
    case 1: return 'G';
    case 2: return 'C';
    default: return 'A'; /* 3 and everything out of range */
    }
}

/* binning.c:114-124 */
int getscore(char *string)
{
    int score = 0;
    for (; *string != '\0'; string++) score = score * 4 + getval(*string);
    return score;
}

static void decode_code(uint64_t hi, uint64_t lo, int n, char *dst)
{
    for (int t = n - 1; t >= 0; t--) {
        dst[t] = getbp((int)(lo & 3));
        lo = (lo >> 2) | (hi << 62);
        hi >>= 2;
    }
    dst[n] = '\0';
}

static void kmer_string(const gbin_table *t, uint64_t s, char *dst)
{
    if (t->kmer_words == 1) decode_code(0, t->kmer_codes[s], t->kmer_size, dst);
    else decode_code(t->kmer_codes[2 * s], t->kmer_codes[2 * s + 1], t->kmer_size, dst);
}

/* ------------------------------------------------------------------ main's read loop */

/* binning.c:1154-1166: `while (fgets(read, READ_LENGTH, file)) { len = strlen(read); read[--len] = 0;
 * process_read(table, read, read_id++); }`.  fgets stores at most READ_LENGTH-1 bytes and stops after
 * a newline; the loop then drops the last byte whatever it is, and every fgets return owns an id. */
int gbin_read_file_fgets(const char *path, int read_length_define, char **data_out, uint64_t *data_bytes_out,
                         uint64_t **starts_out, uint32_t **lens_out, uint64_t *n_reads_out)
{
    if (!path || read_length_define < 2 || !data_out || !starts_out || !lens_out || !n_reads_out) return GBIN_E_INVALID_ARG;
    FILE *f = fopen(path, "rb");
    if (!f) return GBIN_E_IO;
    if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return GBIN_E_IO; }
    long size = ftell(f);
    if (size < 0) { fclose(f); return GBIN_E_IO; }
    rewind(f);
    char *data = malloc((size_t)size + 64);
    if (!data) { fclose(f); return GBIN_E_NOMEM; }
    if (fread(data, 1, (size_t)size, f) != (size_t)size) { fclose(f); free(data); return GBIN_E_IO; }
    fclose(f);
    memset(data + size, 0, 64);

    size_t cap = 1024, n = 0, pos = 0;
    uint64_t *starts = malloc(cap * sizeof *starts);
    uint32_t *lens = malloc(cap * sizeof *lens);
    if (!starts || !lens) { free(data); free(starts); free(lens); return GBIN_E_NOMEM; }
    while (pos < (size_t)size) {
        size_t take = 0;
        while (take < (size_t)(read_length_define - 1) && pos + take < (size_t)size) {
            char c = data[pos + take++];
            if (c == '\n') break;
        }
        if (n == cap) {
            cap *= 2;
            uint64_t *s2 = realloc(starts, cap * sizeof *starts);
            uint32_t *l2 = realloc(lens, cap * sizeof *lens);
            if (!s2 || !l2) { free(data); free(s2 ? s2 : starts); free(l2 ? l2 : lens); return GBIN_E_NOMEM; }
            starts = s2;
            lens = l2;
        }
        starts[n] = pos;
        lens[n] = (uint32_t)(take - 1); /* read[--len] = '\0' */
        n++;
        pos += take;
    }
    *data_out = data;
    if (data_bytes_out) *data_bytes_out = (uint64_t)size;
    *starts_out = starts;
    *lens_out = lens;
    *n_reads_out = n;
    return GBIN_OK;
}

/* ------------------------------------------------------------------ table utilities */

void gbin_table_free(gbin_table *t)
{
    if (!t) return;
    if (!t->ctx_owned && !t->on_device) {
        free(t->mmer_codes);
        free(t->mmer_kmer_off);
        free(t->kmer_codes);
        free(t->kmer_id_off);
        free(t->read_ids);
    }
    t->mmer_codes = NULL;
    t->mmer_kmer_off = NULL;
    t->kmer_codes = NULL;
    t->kmer_id_off = NULL;
    t->read_ids = NULL;
    t->n_buckets = t->n_kmers = t->n_ids = 0;
}

int gbin_table_clone(const gbin_table *h, gbin_table *out)
{
    if (!h || !out || h->on_device) return GBIN_E_INVALID_ARG;
    const uint64_t B = h->n_buckets, S = h->n_kmers, NS = h->n_ids;
    const int KW = h->kmer_words;
    *out = *h;
    out->ctx_owned = 0;
    out->mmer_codes = malloc((B + 1) * sizeof(uint32_t));
    out->mmer_kmer_off = malloc((B + 1) * sizeof(uint64_t));
    out->kmer_codes = malloc((S * KW + 1) * sizeof(uint64_t));
    out->kmer_id_off = malloc((S + 1) * sizeof(uint64_t));
    out->read_ids = malloc((NS + 1) * sizeof(int32_t));
    if (!out->mmer_codes || !out->mmer_kmer_off || !out->kmer_codes || !out->kmer_id_off || !out->read_ids) {
        gbin_table_free(out);
        return GBIN_E_NOMEM;
    }
    memcpy(out->mmer_codes, h->mmer_codes, B * sizeof(uint32_t));
    memcpy(out->mmer_kmer_off, h->mmer_kmer_off, (B + 1) * sizeof(uint64_t));
    memcpy(out->kmer_codes, h->kmer_codes, S * KW * sizeof(uint64_t));
    memcpy(out->kmer_id_off, h->kmer_id_off, (S + 1) * sizeof(uint64_t));
    memcpy(out->read_ids, h->read_ids, NS * sizeof(int32_t));
    return GBIN_OK;
}

static FILE *open_out(const char *path)
{
    if (!path || strcmp(path, "-") == 0) return stdout;
    return fopen(path, "w");
}

int gbin_table_dump(const gbin_table *t, const char *path)
{
    if (!t || t->on_device) return GBIN_E_INVALID_ARG;
    FILE *f = open_out(path);
    if (!f) return GBIN_E_IO;
    char mm[40], km[80];
    for (uint64_t b = 0; b < t->n_buckets; b++) {
        decode_code(0, t->mmer_codes[b], t->mmer_size, mm);
        for (uint64_t s = t->mmer_kmer_off[b]; s < t->mmer_kmer_off[b + 1]; s++) {
            kmer_string(t, s, km);
            fputs(mm, f);
            fputc(' ', f);
            fputs(km, f);
            for (uint64_t q = t->kmer_id_off[s]; q < t->kmer_id_off[s + 1]; q++) fprintf(f, " %d", t->read_ids[q]);
            fputc('\n', f);
        }
    }
    if (f != stdout) fclose(f);
    else fflush(f);
    return GBIN_OK;
}

/* Layout of print_kmer_read_ids (binning.c:792-823) applied to the table as it is right after
 * prune_data, i.e. one id line per k-mer (after expand_read_id_list the reference prints K such lines). */
int gbin_table_dump_reference_format(const gbin_table *t, const char *path)
{
    if (!t || t->on_device) return GBIN_E_INVALID_ARG;
    FILE *f = open_out(path);
    if (!f) return GBIN_E_IO;
    char mm[40], km[80];
    for (uint64_t b = 0; b < t->n_buckets; b++) {
        decode_code(0, t->mmer_codes[b], t->mmer_size, mm);
        fprintf(f, "%s\n", mm);
        for (uint64_t s = t->mmer_kmer_off[b]; s < t->mmer_kmer_off[b + 1]; s++) {
            kmer_string(t, s, km);
            fprintf(f, "%s\n", km);
            for (uint64_t q = t->kmer_id_off[s]; q < t->kmer_id_off[s + 1]; q++) fprintf(f, "%d ", t->read_ids[q]);
            fputc('\n', f);
        }
        fputc('\n', f);
    }
    if (f != stdout) fclose(f);
    else fflush(f);
    return GBIN_OK;
}

/* What print_kmer_read_ids (binning.c:792-823) prints once expand_read_id_list (binning.c:857-888) has given every base of
 * every surviving k-mer its own copy of the k-mer's id list: m-mer line; per k-mer a key line and K lines of ids; a blank
 * line after each bucket.  This is the layout generate_reads.py:14-62 parses. */
int gbin_table_dump_expanded_format(const gbin_table *t, const char *path)
{
    if (!t || t->on_device) return GBIN_E_INVALID_ARG;
    FILE *f = open_out(path);
    if (!f) return GBIN_E_IO;
    char mm[40], km[80];
    for (uint64_t b = 0; b < t->n_buckets; b++) {
        decode_code(0, t->mmer_codes[b], t->mmer_size, mm);
        fprintf(f, "%s\n", mm);
        for (uint64_t s = t->mmer_kmer_off[b]; s < t->mmer_kmer_off[b + 1]; s++) {
            kmer_string(t, s, km);
            fprintf(f, "%s\n", km);
            for (int base = 0; base < t->kmer_size; base++) {
                for (uint64_t q = t->kmer_id_off[s]; q < t->kmer_id_off[s + 1]; q++) fprintf(f, "%d ", t->read_ids[q]);
                fputc('\n', f);
            }
        }
        fputc('\n', f);
    }
    if (f != stdout) fclose(f);
    else fflush(f);
    return GBIN_OK;
}

/* The same text from the expanded lists (gbin_expand_read_ids_device + gbin_expanded_to_host): one line per list of lists entry. */
int gbin_table_dump_expanded_lists(const gbin_table *t, const gbin_expanded *x, const char *path)
{
    if (!t || !x || t->on_device || x->on_device) return GBIN_E_INVALID_ARG;
    if (x->kmer_size != t->kmer_size || x->n_lists != t->n_kmers * (uint64_t)t->kmer_size) return GBIN_E_INVALID_ARG;
    FILE *f = open_out(path);
    if (!f) return GBIN_E_IO;
    char mm[40], km[80];
    for (uint64_t b = 0; b < t->n_buckets; b++) {
        decode_code(0, t->mmer_codes[b], t->mmer_size, mm);
        fprintf(f, "%s\n", mm);
        for (uint64_t s = t->mmer_kmer_off[b]; s < t->mmer_kmer_off[b + 1]; s++) {
            kmer_string(t, s, km);
            fprintf(f, "%s\n", km);
            for (uint64_t l = s * (uint64_t)t->kmer_size; l < (s + 1) * (uint64_t)t->kmer_size; l++) {
                for (uint64_t q = x->list_off[l]; q < x->list_off[l + 1]; q++) fprintf(f, "%d ", x->ids[q]);
                fputc('\n', f);
            }
        }
        fputc('\n', f);
    }
    if (f != stdout) fclose(f);
    else fflush(f);
    return GBIN_OK;
}

void gbin_expanded_free(gbin_expanded *x)
{
    if (!x || x->on_device) return;
    free(x->list_off);
    free(x->ids);
    x->list_off = NULL;
    x->ids = NULL;
    x->n_lists = x->n_ids = 0;
}

/* ------------------------------------------------------------------ ZHashTable / ll_node adapter */

/* zhash.c:171-182: h = (17*h + ch) % size over the key bytes — needed so that the reference's own
 * zhash_get (zhash.c:82-93) finds the entries we place. */
static size_t ref_hash(const char *key, size_t size)
{
    size_t h = 0;
    char ch;
    while ((ch = *key++)) h = (17 * h + (size_t)ch) % size;
    return h;
}

/* zhash.c:75-79: a table grows one step whenever entry_count > size/2 after an insert, so a table
 * that received `count` inserts sits at the first size with count <= size/2. */
static size_t size_index_for(size_t count)
{
    size_t idx = 0;
    while (idx + 1 < GBIN_ZHASH_NUM_SIZES && count > gbin_zhash_sizes[idx] / 2) idx++;
    return idx;
}

static int table_init(struct ZHashTable *t, size_t count)
{
    t->size_index = size_index_for(count);
    t->entry_count = 0;
    t->entries = calloc(gbin_zhash_sizes[t->size_index], sizeof(void *));
    return t->entries ? 0 : -1;
}

/* zcreate_entry + head-of-chain insert (zhash.c:69-73, 148-161): key strcpy-owned, value borrowed */
static int table_put(struct ZHashTable *t, const char *key, void *val)
{
    struct ZHashEntry *e = malloc(sizeof *e);
    char *k = malloc(strlen(key) + 1);
    if (!e || !k) { free(e); free(k); return -1; }
    strcpy(k, key);
    e->key = k;
    e->val = val;
    const size_t h = ref_hash(key, gbin_zhash_sizes[t->size_index]);
    e->next = t->entries[h];
    t->entries[h] = e;
    t->entry_count++;
    return 0;
}

static void free_list(ll_node *n)
{
    while (n) {
        ll_node *nx = n->next;
        free(n);
        n = nx;
    }
}

void gbin_zhash_release(struct ZHashTable *table)
{
    if (!table || !table->entries) return;
    const size_t size = gbin_zhash_sizes[table->size_index];
    for (size_t i = 0; i < size; i++) {
        struct ZHashEntry *e = table->entries[i];
        while (e) {
            struct ZHashEntry *nx = e->next;
            struct ZHashTable *kt = e->val;
            if (kt) {
                const size_t ks = gbin_zhash_sizes[kt->size_index];
                for (size_t j = 0; j < ks; j++) {
                    struct ZHashEntry *ke = kt->entries[j];
                    while (ke) {
                        struct ZHashEntry *knx = ke->next;
                        free_list(ke->val);
                        free(ke->key);
                        free(ke);
                        ke = knx;
                    }
                }
                free(kt->entries);
                free(kt);
            }
            free(e->key);
            free(e);
            e = nx;
        }
    }
    free(table->entries);
    table->entries = NULL;
    table->entry_count = 0;
    table->size_index = 0;
}

int gbin_table_to_zhash(const gbin_table *t, struct ZHashTable *into)
{
    if (!t || !into || t->on_device) return GBIN_E_INVALID_ARG;
    if (into->entries) { /* e.g. the empty 53-slot table from zcreate_hash_table (zhash.c:19-35) */
        if (into->entry_count != 0) return GBIN_E_STATE;
        free(into->entries);
        into->entries = NULL;
    }
    if (table_init(into, t->n_buckets)) return GBIN_E_NOMEM;
    char mm[40], km[80];
    for (uint64_t b = 0; b < t->n_buckets; b++) {
        const uint64_t s0 = t->mmer_kmer_off[b], s1 = t->mmer_kmer_off[b + 1];
        struct ZHashTable *kt = malloc(sizeof *kt);
        if (!kt || table_init(kt, s1 - s0)) { free(kt); return GBIN_E_NOMEM; }
        for (uint64_t s = s0; s < s1; s++) {
            ll_node *head = NULL, **tail = &head;
            for (uint64_t q = t->kmer_id_off[s]; q < t->kmer_id_off[s + 1]; q++) { /* create_node_num, llist.c:6-11 */
                ll_node *n = malloc(sizeof *n);
                if (!n) { free_list(head); return GBIN_E_NOMEM; }
                n->next = NULL;
                n->item = NULL;
                n->read_id = t->read_ids[q];
                *tail = n;
                tail = &n->next;
            }
            kmer_string(t, s, km);
            if (table_put(kt, km, head)) return GBIN_E_NOMEM;
        }
        decode_code(0, t->mmer_codes[b], t->mmer_size, mm);
        if (table_put(into, mm, kt)) return GBIN_E_NOMEM;
    }
    return GBIN_OK;
}

/* ------------------------------------------------------------------ process_read / prune_data */

typedef struct ref_session {
    struct ZHashTable *table; /* identity of the session: the caller's table pointer */
    char *data;
    size_t data_len, data_cap;
    uint64_t *starts;
    uint32_t *lens;
    int32_t *ids;
    size_t n, cap;
    int pruned;
} ref_session;

/* Sessions live in a list that grows on demand; a session ends with prune_data (its slot is free again). */
static ref_session *g_sessions = NULL;
static int g_n_sessions = 0;
static gbin_config g_cfg = {31, 4, 1, 0}; /* binning.c:10-12 defaults */
static gbin_ctx *g_ctx = NULL;
static gbin_config g_ctx_cfg;
static int g_status = GBIN_OK;

int gbin_ref_configure(int kmer_size, int mmer_size, int abundance_cutoff, int device)
{
    gbin_config c = {kmer_size, mmer_size, abundance_cutoff, device};
    if (mmer_size < 2 || mmer_size > 15 || kmer_size < 2 * mmer_size || kmer_size > 64) return GBIN_E_INVALID_CONFIG;
    g_cfg = c;
    return GBIN_OK;
}

int gbin_ref_last_status(void) { return g_status; }

static ref_session *session_for(struct ZHashTable *table, int create)
{
    ref_session *free_slot = NULL;
    for (int i = 0; i < g_n_sessions; i++) {
        if (g_sessions[i].table == table) return &g_sessions[i];
        if (!g_sessions[i].table && !free_slot) free_slot = &g_sessions[i];
    }
    if (!create) return NULL;
    if (!free_slot) {
        const int nn = g_n_sessions ? 2 * g_n_sessions : 8;
        ref_session *ns = realloc(g_sessions, (size_t)nn * sizeof *ns);
        if (!ns) return NULL;
        memset(ns + g_n_sessions, 0, (size_t)(nn - g_n_sessions) * sizeof *ns);
        g_sessions = ns;
        free_slot = &g_sessions[g_n_sessions];
        g_n_sessions = nn;
    }
    memset(free_slot, 0, sizeof *free_slot);
    free_slot->table = table;
    return free_slot;
}

/* The shims cannot return a status (the reference's functions return the table): say on stderr why a call did nothing, once per
 * status change, and keep the status for gbin_ref_last_status(). */
static void shim_fail(int status, const char *what)
{
    static int last_reported = 0;
    g_status = status;
    if (status != last_reported) {
        fprintf(stderr, "libgbin: %s: %s (gbin_ref_last_status() = %d)\n", what, gbin_strerror(status), status);
        last_reported = status;
    }
}

static void session_drop(ref_session *s)
{
    free(s->data);
    free(s->starts);
    free(s->lens);
    free(s->ids);
    memset(s, 0, sizeof *s);
}

/* binning.c:902.  The reference copies what it keeps (keys are strcpy'd, zhash.c:153-157), so the
 * caller may reuse `read` immediately: the read is copied into the session's staging buffer. */
struct ZHashTable *process_read(struct ZHashTable *hash_table, char *read, int read_id)
{
    ref_session *s = session_for(hash_table, 0);
    if (!s) {
        /* A table that already holds entries and has no open session was filled by prune_data: the pointer graph cannot take
         * further inserts.  (An empty table — fresh from zcreate_hash_table, possibly at a recycled address — starts a session.) */
        if (hash_table && hash_table->entry_count != 0) {
            shim_fail(GBIN_E_STATE, "process_read on a table that prune_data has already filled: the read is dropped");
            return hash_table;
        }
        s = session_for(hash_table, 1);
    }
    if (!s) { shim_fail(GBIN_E_NOMEM, "process_read: the read is dropped"); return hash_table; }
    const size_t len = strlen(read);
    if (s->data_len + len + 1 > s->data_cap) {
        size_t nc = s->data_cap ? s->data_cap * 2 : (1u << 20);
        while (nc < s->data_len + len + 1) nc *= 2;
        char *d = realloc(s->data, nc);
        if (!d) { shim_fail(GBIN_E_NOMEM, "process_read: the read is dropped"); return hash_table; }
        s->data = d;
        s->data_cap = nc;
    }
    if (s->n == s->cap) {
        size_t nc = s->cap ? s->cap * 2 : 4096;
        uint64_t *st = realloc(s->starts, nc * sizeof *st);
        if (st) s->starts = st;
        uint32_t *ln = realloc(s->lens, nc * sizeof *ln);
        if (ln) s->lens = ln;
        int32_t *id = realloc(s->ids, nc * sizeof *id);
        if (id) s->ids = id;
        if (!st || !ln || !id) { shim_fail(GBIN_E_NOMEM, "process_read: the read is dropped"); return hash_table; }
        s->cap = nc;
    }
    memcpy(s->data + s->data_len, read, len);
    s->data[s->data_len + len] = '\n';
    s->starts[s->n] = s->data_len;
    s->lens[s->n] = (uint32_t)len;
    s->ids[s->n] = read_id;
    s->n++;
    s->data_len += len + 1;
    g_status = GBIN_OK;
    return hash_table;
}

/* binning.c:1130.  Flush: GPU pipeline over everything staged for this table, then the pointer graph. */
struct ZHashTable *prune_data(struct ZHashTable *hash_table)
{
    ref_session *s = session_for(hash_table, 0);
    if (!s) {
        if (hash_table && hash_table->entry_count != 0) {
            shim_fail(GBIN_E_STATE, "prune_data on a table that it has already filled: nothing done");
            return hash_table;
        }
        s = session_for(hash_table, 1); /* no reads were staged: an empty batch */
    }
    if (!s) { shim_fail(GBIN_E_NOMEM, "prune_data"); return hash_table; }
    if (!g_ctx || memcmp(&g_ctx_cfg, &g_cfg, sizeof g_cfg) != 0) {
        if (g_ctx) gbin_destroy(g_ctx);
        g_ctx = NULL;
        const int rc = gbin_create(&g_cfg, &g_ctx);
        if (rc != GBIN_OK) {
            shim_fail(rc, "prune_data: no GPU context, the table stays empty");
            session_drop(s);
            return hash_table;
        }
        g_ctx_cfg = g_cfg;
    }
    gbin_reads rd;
    memset(&rd, 0, sizeof rd);
    rd.data = s->data;
    rd.data_bytes = s->data_len;
    rd.n_reads = s->n;
    rd.starts = s->starts;
    rd.lens = s->lens;
    rd.read_ids = s->ids;
    gbin_table tab;
    int rc = gbin_bin_reads_host(g_ctx, &rd, &tab);
    if (rc == GBIN_OK) rc = gbin_table_to_zhash(&tab, hash_table);
    if (rc != GBIN_OK) shim_fail(rc, "prune_data: the pipeline failed, the table stays empty");
    else g_status = GBIN_OK;
    /* the session ends here: its slot is free again, whatever address the next table has */
    session_drop(s);
    return hash_table;
}

/* Forget the session bound to a table pointer (lets the same address be reused for a new table). */
void gbin_ref_reset(struct ZHashTable *hash_table)
{
    ref_session *s = session_for(hash_table, 0);
    if (s) session_drop(s);
}
