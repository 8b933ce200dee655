/*
 * gbin_ref_types.h — the reference's result container layouts, restated so that graphs built by
 * libgbin.so are walkable by unmodified reference code.
 *
 *   struct ZHashEntry { char *key; void *val; struct ZHashEntry *next; }            zhash.h:14-18   24 B
 *   struct ZHashTable { size_t size_index; size_t entry_count; ZHashEntry **entries } zhash.h:22-26  24 B
 *   ll_node           { ll_node *next; union { int read_id; void *item; }; }         llist.h:7-13    16 B
 *
 * A translation unit that already includes the reference's zhash.h / llist.h must define
 * GBIN_HAVE_REFERENCE_HEADERS before including this file (the types are then the reference's own).
 */
#ifndef GBIN_REF_TYPES_H
#define GBIN_REF_TYPES_H

#include <stddef.h>

#ifndef GBIN_HAVE_REFERENCE_HEADERS
struct ZHashEntry {
    char *key;
    void *val;
    struct ZHashEntry *next;
};

struct ZHashTable {
    size_t size_index;
    size_t entry_count;
    struct ZHashEntry **entries;
};

typedef struct ll_node {
    struct ll_node *next;
    union {
        int read_id;
        void *item;
    };
} ll_node;
#endif

/* zhash.c:13-17: the prime bucket counts a table steps through (size_index indexes this list). */
#define GBIN_ZHASH_NUM_SIZES 23
static const size_t gbin_zhash_sizes[GBIN_ZHASH_NUM_SIZES] = {
    53, 101, 211, 503, 1553, 3407, 6803, 12503, 25013, 50261, 104729, 250007, 500009, 1000003,
    2000029, 4000037, 10000019, 25000009, 50000047, 104395301, 217645177, 512927357, 1000000007};

#endif
