/* shim_bench.c — times the reference-named entry points of libgbin.so the way the reference's main drives them.
 *
 *   shim_bench <reads-file> K M cutoff READ_LENGTH
 *
 * main's loop (binning.c:1154-1169): fgets a line, drop its last character, process_read(table, line, id++); then prune_data(table).
 * With libgbin.so behind those names process_read only stages the read (host memory); prune_data runs the GPU path and builds the
 * ZHashTable / ZHashEntry / ll_node graph the reference's iterators walk.  Prints one JSON line with the seconds of each phase and
 * the number of k-mer instances, surviving k-mers (graph entries) and id nodes.
 * Build: gcc -O2 tools/shim_bench.c -Iinclude -Lgenome-assembly_b200 -lgbin -Wl,-rpath,$PWD/genome-assembly_b200 -o tools/shim_bench */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gbin.h"
#include "gbin_ref_types.h"

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    if (argc < 6) {
        fprintf(stderr, "usage: %s <reads-file> K M cutoff READ_LENGTH\n", argv[0]);
        return 2;
    }
    const int K = atoi(argv[2]), M = atoi(argv[3]), cutoff = atoi(argv[4]), read_length = atoi(argv[5]);
    if (gbin_ref_configure(K, M, cutoff, 0) != GBIN_OK) {
        fprintf(stderr, "gbin_ref_configure failed\n");
        return 1;
    }
    FILE *f = fopen(argv[1], "r");
    if (!f) {
        perror(argv[1]);
        return 1;
    }
    char *line = malloc((size_t)read_length + 8);
    struct ZHashTable table = {0, 0, NULL};
    int id = 0;
    uint64_t inst = 0;
    /* warm-up of the context (CUDA context creation, first allocations) outside the timed phases: one tiny session */
    {
        struct ZHashTable warm = {0, 0, NULL};
        char tiny[160];
        memset(tiny, 'A', sizeof tiny);
        tiny[K + 4] = 0;
        process_read(&warm, tiny, 0);
        prune_data(&warm);
        gbin_zhash_release(&warm);
        gbin_ref_reset(&warm);
    }
    const double t0 = now();
    while (fgets(line, read_length, f)) {
        const size_t n = strlen(line);
        if (n) line[n - 1] = 0; /* binning.c:1160 */
        if ((int)n - 1 >= K) inst += (uint64_t)((int)n - 1 - K + 1);
        process_read(&table, line, id++);
    }
    const double t1 = now();
    prune_data(&table);
    const double t2 = now();
    const int status = gbin_ref_last_status();
    uint64_t kmers = 0, nodes = 0, buckets = 0;
    if (status == GBIN_OK && table.entries) {
        const size_t size = gbin_zhash_sizes[table.size_index];
        for (size_t i = 0; i < size; i++)
            for (struct ZHashEntry *e = table.entries[i]; e; e = e->next) {
                buckets++;
                const struct ZHashTable *kt = (const struct ZHashTable *)e->val;
                const size_t ks = gbin_zhash_sizes[kt->size_index];
                for (size_t j = 0; j < ks; j++)
                    for (struct ZHashEntry *ke = kt->entries[j]; ke; ke = ke->next) {
                        kmers++;
                        for (const ll_node *p = (const ll_node *)ke->val; p; p = p->next) nodes++;
                    }
            }
    }
    const double t3 = now();
    printf("{\"reads\": %d, \"instances\": %" PRIu64 ", \"status\": %d, \"process_read_s\": %.6f, \"prune_data_s\": %.6f, \"walk_s\": %.6f, "
           "\"buckets\": %" PRIu64 ", \"kmers\": %" PRIu64 ", \"id_nodes\": %" PRIu64 "}\n",
           id, inst, status, t1 - t0, t2 - t1, t3 - t2, buckets, kmers, nodes);
    gbin_zhash_release(&table);
    gbin_ref_reset(&table);
    free(line);
    fclose(f);
    return status == GBIN_OK ? 0 : 1;
}
