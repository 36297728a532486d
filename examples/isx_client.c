/*
 * Minimal C client of the drop-in boundary (include/isx.h -> libisx_b200.so): what a cgo / JNI / N-API binding does,
 * with no Python and no torch in the process. It stores four codes of different lengths, searches one 64-bit query and
 * prints the neighbours. Without a CUDA device isx_open fails, the client prints the library's message and exits 2 -
 * there is no CPU path behind this ABI.
 *
 *   gcc -std=c99 -Iinclude examples/isx_client.c -o /tmp/isx_client -Liscc_search_b200 -lisx_b200 \
 *       -Wl,-rpath,$PWD/iscc_search_b200
 *
 * The search is the one behind /root/reference/iscc_search/indexes/usearch/index.py:2024-2045
 * (`_search_similarity_unit`: `ShardedNphdIndex.search(query, count)` -> keys + NPHD distances, ascending).
 */
#include <stdio.h>
#include <string.h>

#include "isx.h"

static int fail(const char* what, int rc, isx_store_t* store) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, isx_last_error());
    if (store) isx_close(store);
    return store ? 1 : 2;
}

int main(void) {
    isx_store_t* store = NULL;
    int rc = isx_open(&store, 0, 8, 32, 0); /* device 0, uint64 keys, codes up to 32 bytes, variable length (NPHD) */
    if (rc != 0) return fail("isx_open", rc, NULL);

    /* rows of 32 bytes, zero padded; lens gives the code length of each row */
    const uint64_t keys[4] = {11, 12, 13, 14};
    const uint8_t lens[4] = {8, 16, 8, 32};
    uint8_t codes[4][32];
    memset(codes, 0, sizeof codes);
    codes[0][0] = 0xff;  /* key 11: 8 of 64 bits differ from the all-zero query */
    codes[1][15] = 0x01; /* key 12: differs only beyond the 8 bytes a 64-bit query is compared on -> distance 0 */
    codes[2][0] = 0x03;  /* key 13: 2 of 64 bits */
    codes[3][31] = 0x80; /* key 14: likewise distance 0 on the common prefix; ties are ordered by key */
    uint8_t added[4];
    rc = isx_add(store, keys, &codes[0][0], lens, 4, added);
    if (rc != 0) return fail("isx_add", rc, store);

    uint8_t query[32];
    memset(query, 0, sizeof query);
    const uint8_t qlen = 8;
    uint64_t out_keys[4];
    uint16_t out_h[4], out_bits[4];
    uint32_t out_n = 0;
    rc = isx_search(store, query, &qlen, 1, 4, 0, 0, out_keys, out_h, out_bits, &out_n, NULL, NULL);
    if (rc != 0) return fail("isx_search", rc, store);
    for (uint32_t i = 0; i < out_n; i++)
        printf("key %llu  nphd %u/%u\n", (unsigned long long)out_keys[i], (unsigned)out_h[i], (unsigned)out_bits[i]);
    /* by the definition of the metric: 12 0/64, 14 0/64 (ties by key), 13 2/64, 11 8/64 */
    isx_close(store);
    return 0;
}
