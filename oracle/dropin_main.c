/*
 * dropin_main.c — TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * The reference's main() flow (binning.c:1147-1181) with the two hot-path calls bound to libgbin.so instead of the
 * reference's own definitions: process_read / prune_data here are the GPU-backed symbols exported by libgbin.so, while
 * everything downstream — expand_read_id_list, find_kmer_extensions, print_kmers, the iterators, zhash, llist — is the
 * UNMODIFIED reference code (binning.c is compiled by oracle/build_dropin.sh with its own process_read, prune_data,
 * getval, getbp, getscore and main renamed out of the way by -D macros so that the symbols do not clash).
 * This is the link-time substitution INTEGRATION.md describes, exercised end to end.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "zhash.h" /* the reference's own headers (-I/root/reference) */
#include "llist.h"
#define GBIN_HAVE_REFERENCE_HEADERS
#include "gbin.h"

#ifndef DROPIN_READ_LENGTH
#error "build with -DDROPIN_READ_LENGTH / -DDROPIN_K / -DDROPIN_M / -DDROPIN_CUTOFF"
#endif

/* reference functions that stay in use (binning.c:857, 659, 827) */
void expand_read_id_list(struct ZHashTable *hashtable);
void find_kmer_extensions(struct ZHashTable *hash_table, bool left_to_right);
void print_kmers(struct ZHashTable *hash_table);

int main(int argc, char *argv[])
{
    if (argc < 2) return 2;
    if (gbin_ref_configure(DROPIN_K, DROPIN_M, DROPIN_CUTOFF, 0) != GBIN_OK) return 3;
    FILE *file = fopen(argv[1], "r");
    if (!file) return 2;
    struct ZHashTable *hash_table = zcreate_hash_table(); /* the reference's own allocator (zhash.c:19-35) */
    char read[DROPIN_READ_LENGTH];
    int read_id = 0;
    while (fgets(read, DROPIN_READ_LENGTH, file) != NULL) { /* binning.c:1158-1166 verbatim */
        int len = strlen(read);
        read[--len] = '\0';
        process_read(hash_table, read, read_id++);
    }
    prune_data(hash_table);
    if (gbin_ref_last_status() != GBIN_OK) {
        fprintf(stderr, "gbin status %d\n", gbin_ref_last_status());
        return 4;
    }
    if (argc > 2 && strcmp(argv[2], "--table-only") == 0) { /* stop after the hot path: dump what print_kmers sees now */
        print_kmers(hash_table);
        return 0;
    }
    expand_read_id_list(hash_table);
    find_kmer_extensions(hash_table, true);
    find_kmer_extensions(hash_table, false);
    print_kmers(hash_table);
    return 0;
}
