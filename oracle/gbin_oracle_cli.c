/*
 * gbin_oracle_cli.c — TEST INFRASTRUCTURE ONLY.
 * Runs the restated oracle the way oracle/ref_harness_main.c runs the reference:
 *   gbin_oracle_cli <reads-file> K M CUTOFF READ_LENGTH [--strings] [--time]
 * and dumps "<mmer> <kmer> <ids...>" lines, so `sort | md5sum` of both can be compared.
 */
#include "gbin_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s <reads-file> K M CUTOFF READ_LENGTH [--strings] [--time]\n", argv[0]);
        return 2;
    }
    int K = atoi(argv[2]), M = atoi(argv[3]), C = atoi(argv[4]), R = atoi(argv[5]);
    int strings = 0, timing = 0;
    for (int a = 6; a < argc; a++) {
        if (!strcmp(argv[a], "--strings")) strings = 1;
        if (!strcmp(argv[a], "--time")) timing = 1;
    }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *data = malloc((size_t)size + 1);
    if (fread(data, 1, (size_t)size, f) != (size_t)size) { perror("fread"); return 2; }
    fclose(f);
    uint64_t *starts;
    uint32_t *lens;
    size_t n = orc_fgets_split(data, (size_t)size, R, &starts, &lens);
    if (strings) return orc_dump_strings(data, starts, lens, n, NULL, K, M, C, stdout);
    orc_result res;
    double t0 = now_s();
    int rc = orc_run(data, starts, lens, n, NULL, K, M, C, &res);
    double t1 = now_s();
    if (rc) { fprintf(stderr, "orc_run failed: %d\n", rc); return 1; }
    if (!timing) orc_dump(&res, stdout);
    fprintf(timing ? stdout : stderr,
            "{\"k\": %d, \"m\": %d, \"cutoff\": %d, \"read_ids\": %zu, \"instances\": %llu, \"distinct\": %llu, "
            "\"surviving_kmers\": %llu, \"surviving_buckets\": %llu, \"ids\": %llu, \"total_s\": %.6f}\n",
            K, M, C, n, (unsigned long long)res.n_instances, (unsigned long long)res.n_distinct,
            (unsigned long long)res.n_kmers, (unsigned long long)res.n_buckets, (unsigned long long)res.n_ids, t1 - t0);
    return 0;
}
