/* CPU ORACLE (C part) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE. PARITY UNPINNED (see scn_oracle.py header).
 *
 * Restates the hash-map algorithm of [UPSTREAM] facebookresearch/SparseConvNet `SCN/Metadata/`:
 *   IOLayersRules.h::inputLayerRules (mode 4), SubmanifoldConvolutionRules.h, ConvolutionRules.h (k2/s2),
 * which the reference reaches only through mopa/models/scn_unet.py:25-30. Upstream keeps one
 * google::dense_hash_map<Point<3>, Int> per batch sample; here one open-addressing table (quadratic
 * probing, load <= 0.5, like dense_hash_map) keyed by the packed (batch, x, y, z). Serial, as upstream's
 * per-sample builders are. Used (a) to cross-check the numpy restatement and (b) as the rulebook leg of
 * the timed CPU baseline in bench.py.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t *keys;
    int32_t *vals;
    uint64_t mask;
} grid_t;

#define EMPTY_KEY 0xFFFFFFFFFFFFFFFFull

static uint64_t pack(const int64_t *c) { /* c = x, y, z, b */
    return ((uint64_t)c[3] << 48) | ((uint64_t)c[0] << 32) | ((uint64_t)c[1] << 16) | (uint64_t)c[2];
}

static uint64_t mix(uint64_t h) { /* splitmix64 finaliser */
    h ^= h >> 30; h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 27; h *= 0x94d049bb133111ebull;
    h ^= h >> 31;
    return h;
}

static int grid_init(grid_t *g, int64_t n) {
    uint64_t cap = 16;
    while (cap < (uint64_t)(2 * n + 2)) cap <<= 1;
    g->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    g->vals = (int32_t *)malloc(cap * sizeof(int32_t));
    if (!g->keys || !g->vals) return -1;
    memset(g->keys, 0xFF, cap * sizeof(uint64_t));
    g->mask = cap - 1;
    return 0;
}

static void grid_free(grid_t *g) { free(g->keys); free(g->vals); }

/* returns the slot of key, or of the empty slot where it would go */
static uint64_t grid_slot(const grid_t *g, uint64_t key) {
    uint64_t s = mix(key) & g->mask, step = 0;
    while (g->keys[s] != EMPTY_KEY && g->keys[s] != key) { step++; s = (s + step) & g->mask; }
    return s;
}

static int32_t grid_find(const grid_t *g, uint64_t key) {
    uint64_t s = grid_slot(g, key);
    return g->keys[s] == key ? g->vals[s] : -1;
}

/* find-or-insert with the running counter; returns the id */
static int32_t grid_get(grid_t *g, uint64_t key, int32_t *ctr) {
    uint64_t s = grid_slot(g, key);
    if (g->keys[s] != key) { g->keys[s] = key; g->vals[s] = (*ctr)++; }
    return g->vals[s];
}

/* inputLayerRules, mode 4: first occurrence of a site gets the next id.
 * coords: (n, ncols) int64, ncols 3 or 4 (batch last). Outputs: p2v (n), voxel_coords (<= n, 4). Returns V. */
int64_t oracle_input_rules(const int64_t *coords, int64_t n, int ncols, int32_t *p2v, int64_t *voxel_coords) {
    grid_t g;
    if (grid_init(&g, n)) return -1;
    int32_t ctr = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t c[4] = {coords[i * ncols], coords[i * ncols + 1], coords[i * ncols + 2],
                        ncols == 4 ? coords[i * ncols + 3] : 0};
        int32_t before = ctr;
        int32_t id = grid_get(&g, pack(c), &ctr);
        if (ctr != before) memcpy(voxel_coords + 4 * (int64_t)id, c, sizeof c);
        p2v[i] = id;
    }
    grid_free(&g);
    return ctr;
}

/* Submanifold 3x3x3 rules as the dense table nbr[k * V + o] = id of site at coord(o) + delta_k, or -1. */
int oracle_subm_rules(const int64_t *voxel_coords, int64_t v, int64_t spatial_size, int32_t *nbr) {
    grid_t g;
    if (grid_init(&g, v)) return -1;
    int32_t ctr = 0;
    for (int64_t i = 0; i < v; i++) grid_get(&g, pack(voxel_coords + 4 * i), &ctr);
    for (int64_t o = 0; o < v; o++) {
        const int64_t *p = voxel_coords + 4 * o;
        int k = 0;
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++)
                for (int dz = -1; dz <= 1; dz++, k++) {
                    int64_t q[4] = {p[0] + dx, p[1] + dy, p[2] + dz, p[3]};
                    int32_t id = -1;
                    if (q[0] >= 0 && q[1] >= 0 && q[2] >= 0 && q[0] < spatial_size && q[1] < spatial_size &&
                        q[2] < spatial_size)
                        id = grid_find(&g, pack(q));
                    nbr[(int64_t)k * v + o] = id;
                }
    }
    grid_free(&g);
    return 0;
}

/* Convolution(k2, s2) rules: parent = coord >> 1, k = (x&1)*4 + (y&1)*2 + (z&1); coarse ids canonical
 * (first occurrence scanning fine ids ascending). Returns Vc. */
int64_t oracle_strided_rules(const int64_t *fine_coords, int64_t vf, int64_t *coarse_coords, int32_t *parent,
                             int32_t *kidx) {
    grid_t g;
    if (grid_init(&g, vf)) return -1;
    int32_t ctr = 0;
    for (int64_t i = 0; i < vf; i++) {
        const int64_t *p = fine_coords + 4 * i;
        int64_t c[4] = {p[0] >> 1, p[1] >> 1, p[2] >> 1, p[3]};
        int32_t before = ctr;
        int32_t id = grid_get(&g, pack(c), &ctr);
        if (ctr != before) memcpy(coarse_coords + 4 * (int64_t)id, c, sizeof c);
        parent[i] = id;
        kidx[i] = (int32_t)((p[0] & 1) * 4 + (p[1] & 1) * 2 + (p[2] & 1));
    }
    grid_free(&g);
    return ctr;
}
