/*
 * gbin_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the k-mer binning hot path of twitu/genome-assembly
 * (binning.c: process_read, prune_data/prune_kmers, getval/getbp/getscore and main's fgets loop).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * build, link, load or execute anything under oracle/.  The product (libgbin.so) never does.
 *
 * Parity status: PINNED by execution of the reference itself (oracle/_ref, built by
 * oracle/build_ref.sh from /root/reference) — the reference ships no tests or golden vectors of
 * its own, so the pins are md5 digests of the reference binary's sorted table
 * dump, committed under tests/golden/ together with the script that made them.
 */
#ifndef GBIN_ORACLE_H
#define GBIN_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* binning.c:91-111, 69-88, 114-124 */
int orc_getval(char c);
char orc_getbp(int bp);
int orc_getscore(const char *string);

/* One k-mer instance (one window of one read), as process_read would insert it
 * (binning.c:1023-1069): m-mer bucket, oriented k-mer, arrival index of the read. */
typedef struct orc_tuple {
    uint32_t mmer;    /* base-4 code of the (possibly complemented) signature, T,G,C,A = 0..3, first char most significant */
    uint32_t arrival; /* index of the read in arrival order (0-based) */
    uint64_t khi;     /* 2K-bit code of the (possibly complemented) k-mer: high 64 bits (0 when K <= 32) */
    uint64_t klo;     /* low 64 bits */
} orc_tuple;

/* Per-window trace of process_read's signature machinery, for kernel-level parity. */
typedef struct orc_window {
    int32_t sig_pos; /* position of the signature m-mer in the read (signature - read, binning.c:952,986,1019) */
    int32_t is_rev;  /* binning.c:943,948 */
    uint32_t mmer;   /* max_score == code of the stored m-mer key */
} orc_window;

/* Restatement of process_read's loop (binning.c:918-1073) for one read of length len.
 * tuples/windows may be NULL; returns the number of windows (max(0, len-K+1)).
 * Non-ACGT bytes are scored as 3 (getval default) — code mode cannot represent the raw byte the
 * reference would keep in an un-flipped key, see orc_dump_strings for that. */
size_t orc_process_read(const char *read, int len, int K, int M, uint32_t arrival,
                        orc_tuple *tuples, orc_window *windows);

/* Replay of main's read loop (binning.c:1154-1166) over a file image: every fgets(buf, READ_LENGTH)
 * return yields one read (its last char dropped) and consumes one read id.
 * Writes n+1 offsets / n lengths into freshly malloc'ed arrays (caller frees). starts[i] is the byte
 * offset of read i in `data`, lens[i] its length after the chop. Returns n. */
size_t orc_fgets_split(const char *data, size_t size, int read_length_define,
                       uint64_t **starts_out, uint32_t **lens_out);

/* Flat result table: the pruned two-level mmer -> kmer -> read-id-list store in canonical order
 * (m-mer code ascending, k-mer code ascending, ids newest first == reverse arrival order). */
typedef struct orc_result {
    int K, M, cutoff, kw;      /* kw = 64-bit words per k-mer code: 1 (K<=32) or 2 */
    uint64_t n_instances;      /* windows processed */
    uint64_t n_distinct;       /* distinct (mmer,kmer) before prune */
    uint64_t n_buckets;        /* surviving m-mer buckets */
    uint64_t n_kmers;          /* surviving k-mers */
    uint64_t n_ids;            /* read-id nodes in surviving lists */
    uint32_t *mmer_codes;      /* [n_buckets] */
    uint64_t *mmer_kmer_off;   /* [n_buckets+1] */
    uint64_t *kmer_codes;      /* [n_kmers*kw], most significant word first */
    uint64_t *kmer_id_off;     /* [n_kmers+1] */
    int32_t *read_ids;         /* [n_ids] */
} orc_result;

/* Whole hot path: process_read over n reads (read i = data[starts[i] .. starts[i]+lens[i]), id
 * ids ? ids[i] : i) followed by prune_data with `cutoff` (keep iff count > cutoff,
 * binning.c:1094-1102).  cutoff < 0 skips pruning. Returns 0 on success. */
int orc_run(const char *data, const uint64_t *starts, const uint32_t *lens, size_t n_reads,
            const int32_t *ids, int K, int M, int cutoff, orc_result *out);
void orc_result_free(orc_result *r);
/* Grouping + prune over externally supplied tuples in arrival order (sorted in place); used by the
 * multi-rank host-logic tests where each rank groups the tuples it received. */
int orc_group_tuples(orc_tuple *t, size_t n, const int32_t *ids, int32_t id_base, int K, int M, int cutoff, orc_result *out);

/* Emits all tuples of all reads in arrival order (pre-sort), for scan-kernel parity.
 * tuples must hold sum(max(0,len-K+1)) entries. Returns the count. */
size_t orc_scan_all(const char *data, const uint64_t *starts, const uint32_t *lens, size_t n_reads,
                    int K, int M, orc_tuple *tuples, orc_window *windows);

/* "<mmer> <kmer> <id> <id> ...\n" per surviving k-mer (same format as oracle/ref_harness_main.c). */
void orc_decode(uint64_t hi, uint64_t lo, int n, char *dst); /* n bases, NUL-terminated */
int orc_dump(const orc_result *r, FILE *f);

/* String-faithful variant (keeps raw bytes in un-flipped keys, binning.c:1023-1040): builds, sorts
 * and prunes on the actual key strings and dumps in the same line format. Used to pin the
 * restatement against the reference on inputs with non-ACGT bytes. */
int orc_dump_strings(const char *data, const uint64_t *starts, const uint32_t *lens, size_t n_reads,
                     const int32_t *ids, int K, int M, int cutoff, FILE *f);

#ifdef __cplusplus
}
#endif
#endif
