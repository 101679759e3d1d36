/*
 * mqvs_demo.c -- the C ABI of libmqvs.so from plain C (no Python, no torch): what a compiled host would bind.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/mqvs_demo.c -Lmediquery-rag_b200 -lmqvs \
 *       -Wl,-rpath,$PWD/mediquery-rag_b200 -lm -o /tmp/mqvs_demo && /tmp/mqvs_demo
 *
 * Does the work of the reference's three call sites (/root/reference/src/ingest_medical.py:106,
 * src/medical_engine.py:52, src/agents/nodes.py:93) on synthetic vectors: create a collection, add rows (K1 normalises
 * them), search with k = 5, print ids and cosine scores.  Without a B200 it prints the library's error and exits 2.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "mqvs.h"

int main(void) {
    const int dim = 768, n = 154, k = 5;
    int rc, n_dev = 0, j, c;
    mqvs_t* store = NULL;
    float *rows, *q, scores[5];
    int64_t ids[5];
    uint64_t first = 0, count = 0;

    printf("libmqvs version %d\n", mqvs_version());
    rc = mqvs_device_count(&n_dev);
    if (rc != MQVS_OK || n_dev < 1) {
        printf("no usable GPU (rc=%d): %s\n", rc, mqvs_last_error());
        return 2;
    }
    rc = mqvs_create(1, NULL, dim, MQVS_F32, 0, &store);
    if (rc != MQVS_OK) {
        printf("mqvs_create failed (rc=%d): %s\n", rc, mqvs_last_error());
        return 2;
    }
    rows = (float*)malloc(sizeof(float) * (size_t)n * dim);
    q = (float*)malloc(sizeof(float) * dim);
    if (!rows || !q) return 1;
    for (j = 0; j < n; ++j)
        for (c = 0; c < dim; ++c) rows[(size_t)j * dim + c] = (float)sin(0.37 * (j + 1) * (c + 1)) + 0.01f * (float)j;
    for (c = 0; c < dim; ++c) q[c] = rows[(size_t)42 * dim + c] + 0.05f * (float)cos(1.3 * c);  /* near row 42 */

    rc = mqvs_add(store, rows, MQVS_F32, 0, (uint64_t)n, &first);
    if (rc == MQVS_OK) rc = mqvs_count(store, &count);
    if (rc == MQVS_OK) rc = mqvs_search(store, q, 0, 1, k, NULL, 0, scores, ids, 0);
    if (rc != MQVS_OK) {
        printf("error (rc=%d): %s\n", rc, mqvs_last_error());
        mqvs_destroy(store);
        return 1;
    }
    printf("%llu rows stored, first row id %llu; top-%d of the query:\n", (unsigned long long)count, (unsigned long long)first, k);
    for (j = 0; j < k; ++j) printf("  row %lld  cosine %.6f\n", (long long)ids[j], scores[j]);
    mqvs_destroy(store);
    free(rows);
    free(q);
    return ids[0] == 42 ? 0 : 1;
}
