/* CPU oracle: affinity-keyed priority flood.   TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's numba kernel
 *   raveled_affinity_watershed   (src/iterseg/watershed.py:95-159)
 * with its heap element Element(value, age, index, source) (watershed.py:162).
 *
 * Semantics restated (line numbers are watershed.py):
 *   - every seed is pushed with value 0.0f and age 0 (:125-131); tuple order
 *     therefore falls through to `index` among seeds;
 *   - pop the smallest (value, age, index) (:134); for the six neighbours in
 *     the order of `offsets` (:135-139): skip if not in mask (:140-142), skip if
 *     already labelled (:143-146), otherwise label it NOW with the popped
 *     voxel's label (:149), key it with the single edge affinity
 *     image[axis_i, aff_off_i + popped] (:150-151) where aff_off_i is 0 for the
 *     first half of the offsets (:119-120) and the neighbour offset for the
 *     second half, bump the global age (:152) and push (:153-154).
 *   - (value, age) with unique age>0 is a strict total order, so any correct
 *     min-heap pops in the same order as Python's heapq.
 *
 * Pinned: bit-identical to the numba kernel and its py_func on the scenes in
 * tests/golden/ (made by scripts/make_golden.py from the verbatim reference).
 *
 * Unlike the reference (which indexes without bounds checks and relies on the
 * zero-padded mask border, watershed.py:213), neighbours outside [0, npix) are
 * treated as outside the mask.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float value;
    int64_t age;
    int64_t index;
} elem_t;

static inline int elem_less(const elem_t *a, const elem_t *b) {
    if (a->value < b->value) return 1;
    if (b->value < a->value) return 0;
    if (a->age != b->age) return a->age < b->age;
    return a->index < b->index;
}

typedef struct {
    elem_t *d;
    int64_t n, cap;
} heap_t;

static int heap_push(heap_t *h, elem_t e) {
    if (h->n == h->cap) {
        int64_t nc = h->cap ? h->cap * 2 : 1024;
        elem_t *nd = (elem_t *)realloc(h->d, (size_t)nc * sizeof(elem_t));
        if (!nd) return -1;
        h->d = nd;
        h->cap = nc;
    }
    int64_t i = h->n++;
    while (i > 0) {
        int64_t p = (i - 1) >> 1;
        if (!elem_less(&e, &h->d[p])) break;
        h->d[i] = h->d[p];
        i = p;
    }
    h->d[i] = e;
    return 0;
}

static elem_t heap_pop(heap_t *h) {
    elem_t top = h->d[0];
    elem_t last = h->d[--h->n];
    int64_t i = 0, n = h->n;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= n) break;
        if (c + 1 < n && elem_less(&h->d[c + 1], &h->d[c])) c++;
        if (!elem_less(&h->d[c], &last)) break;
        h->d[i] = h->d[c];
        i = c;
    }
    if (n > 0) h->d[i] = last;
    return top;
}

/* image: (nchan, npix) float32 C-contiguous; seeds: flat indices (already
 * labelled 1..n in `output` by the caller, watershed.py:61-62); offsets:
 * (n_neighbors, 2) int64 rows (axis, flat offset) (watershed.py:84-92);
 * mask: npix bytes; output: npix uint32, in place.
 * Returns the number of pushes (final age), or -1 on allocation failure. */
int64_t isg_oracle_flood(const float *image, int64_t npix,
                         const int64_t *seeds, int64_t nseeds,
                         const int64_t *offsets, int64_t n_neighbors,
                         const uint8_t *mask, uint32_t *output) {
    heap_t h = {0, 0, 0};
    int64_t age = 0;
    for (int64_t i = 0; i < nseeds; i++) {
        elem_t e = {0.0f, 0, seeds[i]};
        if (heap_push(&h, e)) { free(h.d); return -1; }
    }
    while (h.n > 0) {
        elem_t e = heap_pop(&h);
        for (int64_t i = 0; i < n_neighbors; i++) {
            int64_t axis = offsets[2 * i];
            int64_t off = offsets[2 * i + 1];
            int64_t nb = e.index + off;
            if (nb < 0 || nb >= npix) continue;
            if (!mask[nb]) continue;
            if (output[nb]) continue;
            output[nb] = output[e.index];
            int64_t aoff = (i < n_neighbors / 2) ? 0 : off;
            elem_t ne = {image[axis * npix + aoff + e.index], ++age, nb};
            if (heap_push(&h, ne)) { free(h.d); return -1; }
        }
    }
    free(h.d);
    return age;
}

/* ---------------------------------------------------------------------------
 * Node-keyed priority flood: the classic marker watershed the DoG blob path calls
 * (skimage.segmentation.watershed(image, markers, mask=mask), connectivity 1, no
 * compactness; reference call site src/iterseg/segmentation.py:646).
 *
 * scikit-image (not available here; restated from _watershed_cy.pyx::watershed_raveled):
 *   every marker voxel is pushed in raveled order with value image[index], age 0; pop the
 *   smallest (value, age); for the neighbours in the order of `offsets`: skip if not in
 *   mask or already labelled, else age += 1, label it with the popped voxel's label and
 *   push it with value image[neighbour] and that age.
 * PARITY UNPINNED in one respect: markers of equal value all carry age 0, and the order
 * in which scikit-image's binary heap pops them is an artefact of its sift routines.
 * Here (and in the CUDA kernel) equal-valued markers pop in raveled-index order.
 * key: int64 values, smaller pops first (the caller maps -distance to them).
 */
typedef struct {
    int64_t value;
    int64_t age;
    int64_t index;
} nelem_t;

static inline int nelem_less(const nelem_t *a, const nelem_t *b) {
    if (a->value != b->value) return a->value < b->value;
    if (a->age != b->age) return a->age < b->age;
    return a->index < b->index;
}

int64_t isg_oracle_node_flood(const int64_t *image, int64_t npix, const int64_t *offsets,
                              int64_t n_offsets, const uint8_t *mask, int32_t *output) {
    nelem_t *h = NULL;
    int64_t n = 0, cap = 0, age = 0;
#define NPUSH(e)                                                                      \
    do {                                                                              \
        if (n == cap) {                                                               \
            cap = cap ? cap * 2 : 1024;                                               \
            nelem_t *nd = (nelem_t *)realloc(h, (size_t)cap * sizeof(nelem_t));       \
            if (!nd) { free(h); return -1; }                                          \
            h = nd;                                                                   \
        }                                                                             \
        int64_t i_ = n++;                                                             \
        while (i_ > 0) {                                                              \
            int64_t p_ = (i_ - 1) >> 1;                                               \
            if (!nelem_less(&(e), &h[p_])) break;                                     \
            h[i_] = h[p_];                                                            \
            i_ = p_;                                                                  \
        }                                                                             \
        h[i_] = (e);                                                                  \
    } while (0)
    for (int64_t v = 0; v < npix; ++v)
        if (output[v]) {
            nelem_t e = {image[v], 0, v};
            NPUSH(e);
        }
    while (n > 0) {
        nelem_t top = h[0];
        nelem_t last = h[--n];
        int64_t i = 0;
        for (;;) {
            int64_t c = 2 * i + 1;
            if (c >= n) break;
            if (c + 1 < n && nelem_less(&h[c + 1], &h[c])) c++;
            if (!nelem_less(&h[c], &last)) break;
            h[i] = h[c];
            i = c;
        }
        if (n > 0) h[i] = last;
        for (int64_t k = 0; k < n_offsets; ++k) {
            int64_t nb = top.index + offsets[k];
            if (nb < 0 || nb >= npix || !mask[nb] || output[nb]) continue;
            ++age;
            output[nb] = output[top.index];
            nelem_t e = {image[nb], age, nb};
            NPUSH(e);
        }
    }
#undef NPUSH
    free(h);
    return age;
}
