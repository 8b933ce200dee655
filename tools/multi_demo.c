/* multi_demo.c — the hot path on every GPU of the node from plain C: no Python, no MPI, no NCCL.
 *
 *   multi_demo <reads-file> K M cutoff READ_LENGTH [n_gpus]
 *
 * Replays main's read loop (binning.c:1154-1166) with gbin_read_file_fgets, bins the reads on n_gpus GPUs (default: as many as
 * gbin_multi_create accepts, trying 8, 4, 2, 1) with gbin_multi_bin_reads_host and prints one JSON line: per-GPU table sizes and the
 * sum of the owners' gbin_table_digest — which equals the digest of the single-GPU table (tests/golden/pins.json holds it).
 * Build: gcc -O2 tools/multi_demo.c -Iinclude -Lgenome-assembly_b200 -lgbin -Wl,-rpath,$PWD/genome-assembly_b200 -o tools/multi_demo */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gbin.h"

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s <reads-file> K M cutoff READ_LENGTH [n_gpus]\n", argv[0]);
        return 2;
    }
    gbin_config cfg = {atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), 0};
    const int read_length = atoi(argv[5]);
    char *data = NULL;
    uint64_t bytes = 0, n = 0, *starts = NULL;
    uint32_t *lens = NULL;
    int rc = gbin_read_file_fgets(argv[1], read_length, &data, &bytes, &starts, &lens, &n);
    if (rc) {
        fprintf(stderr, "gbin_read_file_fgets: %s\n", gbin_strerror(rc));
        return 1;
    }
    gbin_multi *m = NULL;
    int G = argc > 6 ? atoi(argv[6]) : 0;
    if (G > 0) rc = gbin_multi_create(&cfg, NULL, G, &m);
    else
        for (G = 8; G >= 1; G /= 2)
            if ((rc = gbin_multi_create(&cfg, NULL, G, &m)) == GBIN_OK) break;
    if (rc || !m) {
        fprintf(stderr, "gbin_multi_create: %s\n", gbin_strerror(rc));
        return 1;
    }
    gbin_reads rd;
    memset(&rd, 0, sizeof rd);
    rd.data = data;
    rd.data_bytes = bytes;
    rd.n_reads = n;
    rd.starts = starts;
    rd.lens = lens;
    gbin_table *t = calloc((size_t)G, sizeof *t);
    for (int rep = 0; rep < 2; rep++) { /* twice: the second call reuses every buffer, exchange epochs advance */
        rc = gbin_multi_bin_reads_host(m, &rd, t);
        if (rc) {
            fprintf(stderr, "gbin_multi_bin_reads_host: %s (%s)\n", gbin_strerror(rc), gbin_multi_last_error(m));
            return 1;
        }
    }
    uint64_t digest = 0, kmers = 0, ids = 0, inst = 0, buckets = 0;
    printf("{\"gpus\": %d, \"reads\": %" PRIu64 ", \"per_gpu_kmers\": [", G, n);
    for (int g = 0; g < G; g++) {
        uint64_t d = 0;
        gbin_table_digest(NULL, &t[g], NULL, &d);
        digest += d;
        kmers += t[g].n_kmers;
        ids += t[g].n_ids;
        inst += t[g].n_instances;
        buckets += t[g].n_buckets;
        printf("%s%" PRIu64, g ? ", " : "", t[g].n_kmers);
    }
    printf("], \"instances\": %" PRIu64 ", \"surviving_kmers\": %" PRIu64 ", \"surviving_ids\": %" PRIu64 ", \"buckets\": %" PRIu64
           ", \"digest\": \"%016" PRIx64 "\"}\n",
           inst, kmers, ids, buckets, digest);
    gbin_multi_destroy(m);
    free(t);
    free(data);
    free(starts);
    free(lens);
    return 0;
}
