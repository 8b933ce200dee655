/*
 * TEST INFRASTRUCTURE ONLY (oracle/): driver appended to the reference translation unit.
 *
 * oracle/build_ref.sh streams /root/reference/binning.c through sed (patching the four
 * unguarded #defines at binning.c:10-13 and, for M > 8, the power_val table at binning.c:17)
 * straight into gcc with -Dmain=ref_main, and appends this file to the stream.  No reference
 * source is copied into the repository; the only outputs are binaries in oracle/_ref/.
 *
 * The driver below repeats the reference's own read loop (binning.c:1150-1169: fopen, fgets
 * into char read[READ_LENGTH], drop the last char, process_read(table, read, read_id++),
 * prune_data) and then walks the two-level table with the reference's own iterators
 * (pattern of print_kmer_read_ids, binning.c:798-819) printing one line per surviving k-mer:
 *     "<mmer> <kmer> <id> <id> ...\n"
 * Modes:
 *   ref_harness <reads-file>            dump to stdout, timings to stderr
 *   ref_harness <reads-file> --time     no dump; prints JSON timing line to stdout
 *   ref_harness <reads-file> --noprune  dump the table before prune_data
 *   ref_harness <reads-file> --expanded the reference's own expand_read_id_list (binning.c:857-888) followed by its own
 *                                       print_kmer_read_ids (binning.c:792-823) on the pruned table: the layout
 *                                       generate_reads.py:14-62 parses (K id lines per k-mer)
 */
#undef main
#include <time.h>

static double hz_now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char *argv[])
{
    int want_dump = 1, want_prune = 1, want_expanded = 0;
    if (argc < 2) {
        fprintf(stderr, "usage: %s <reads-file> [--time|--noprune]\n", argv[0]);
        return 2;
    }
    for (int a = 2; a < argc; a++) {
        if (strcmp(argv[a], "--time") == 0) want_dump = 0;
        if (strcmp(argv[a], "--noprune") == 0) want_prune = 0;
        if (strcmp(argv[a], "--expanded") == 0) want_expanded = 1;
    }
    FILE *file = fopen(argv[1], "r");
    if (!file) { perror(argv[1]); return 2; }

    struct ZHashTable *hash_table = zcreate_hash_table();
    char read[READ_LENGTH];
    int read_id = 0;
    long long instances = 0;

    double t0 = hz_now();
    while (fgets(read, READ_LENGTH, file) != NULL) {
        int len = strlen(read);
        read[--len] = '\0';
        if (len >= KMER_SIZE) instances += len - KMER_SIZE + 1;
        process_read(hash_table, read, read_id++);
    }
    double t1 = hz_now();
    if (want_prune) prune_data(hash_table);
    double t2 = hz_now();
    fclose(file);
    if (want_expanded) {
        expand_read_id_list(hash_table);
        print_kmer_read_ids(hash_table);
        return 0;
    }

    long long survivors = 0, buckets = 0;
    struct ZHashEntry *mmer_entry, *kmer_entry;
    while ((mmer_entry = (struct ZHashEntry *)iterate_level_one_hash(hash_table, false, false)) != NULL) {
        struct ZHashTable *kmer_hash = mmer_entry->val;
        buckets++;
        while ((kmer_entry = (struct ZHashEntry *)iterate_level_two_hash(kmer_hash, false, false)) != NULL) {
            survivors++;
            if (want_dump) {
                fputs(mmer_entry->key, stdout);
                fputc(' ', stdout);
                fputs(kmer_entry->key, stdout);
                for (ll_node *n = (ll_node *)kmer_entry->val; n != NULL; n = n->next)
                    printf(" %d", n->read_id);
                fputc('\n', stdout);
            }
        }
    }
    FILE *rep = want_dump ? stderr : stdout;
    fprintf(rep,
            "{\"k\": %d, \"m\": %d, \"cutoff\": %d, \"read_length_define\": %d, \"read_ids\": %d, "
            "\"instances\": %lld, \"surviving_kmers\": %lld, \"surviving_buckets\": %lld, "
            "\"process_s\": %.6f, \"prune_s\": %.6f}\n",
            KMER_SIZE, MMER_SIZE, ABUNDANCE_CUTOFF, READ_LENGTH, read_id, instances, survivors, buckets,
            t1 - t0, t2 - t1);
    return 0;
}
