/*
 * ORACLE — test infrastructure, not product code.
 *
 * Plain-C restatement of the reference's per-measure store,
 * /root/reference/src/store/in-memory.js:7-431, used (a) by tests/ to check the CUDA
 * path at sizes the pure-Python oracle (oracle/store_oracle.py) cannot reach and
 * (b) by bench.py as the CPU baseline ("port": the reference's algorithm on one host
 * core; the reference itself is JavaScript and no JS engine exists in this image).
 * Nothing in the product package links, loads or calls this file.
 *
 * The reference keeps cells in a JavaScript Map<int, double>: insertion-ordered, `set`
 * on an existing key keeps its slot, `delete` + `set` appends.  `omap` below has
 * exactly those semantics (entry array in insertion order + open-addressing index).
 * Each function follows the reference loop it cites, statement by statement, with
 * JS number semantics (IEEE doubles, Math.max/min NaN-propagating, % = fmod, the
 * Uint16 wrap of `contributions`).
 *
 * Pinned: tests/test_oracle_c.py checks this file against oracle/store_oracle.py,
 * which is itself pinned on the reference's own test expectations (tests/kats.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define O_SUM 0
#define O_AVERAGE 1
#define O_HIGHEST 2
#define O_LOWEST 3
#define O_FIRST 4
#define O_LAST 5
#define O_PRODUCT 6

/* ------------------------------------------------------------------ ordered map */
typedef struct {
    int64_t *keys;  /* insertion order; -1 = deleted entry */
    double *vals;
    int64_t n_entries, n_live, cap_entries;
    int64_t *slots; /* -1 empty, -2 tombstone, else entry index */
    int64_t n_slots, n_used; /* n_used counts non-empty slots (live + tombstones) */
} omap;

static uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static void omap_init(omap *m, int64_t hint) {
    int64_t slots = 16;
    while (slots < hint * 2) slots <<= 1;
    m->cap_entries = hint > 8 ? hint : 8;
    m->keys = (int64_t *)malloc(sizeof(int64_t) * m->cap_entries);
    m->vals = (double *)malloc(sizeof(double) * m->cap_entries);
    m->n_entries = m->n_live = 0;
    m->n_slots = slots;
    m->n_used = 0;
    m->slots = (int64_t *)malloc(sizeof(int64_t) * slots);
    for (int64_t i = 0; i < slots; ++i) m->slots[i] = -1;
}

static void omap_free(omap *m) { free(m->keys); free(m->vals); free(m->slots); }

static void omap_rebuild(omap *m, int64_t min_live) {
    /* compact deleted entries (order kept) and resize the index */
    int64_t w = 0;
    for (int64_t r = 0; r < m->n_entries; ++r)
        if (m->keys[r] >= 0) { m->keys[w] = m->keys[r]; m->vals[w] = m->vals[r]; ++w; }
    m->n_entries = w;
    int64_t want = w > min_live ? w : min_live;
    if (m->cap_entries < want * 2) {
        m->cap_entries = want * 2 + 8;
        m->keys = (int64_t *)realloc(m->keys, sizeof(int64_t) * m->cap_entries);
        m->vals = (double *)realloc(m->vals, sizeof(double) * m->cap_entries);
    }
    int64_t slots = 16;
    while (slots < want * 4) slots <<= 1;
    free(m->slots);
    m->slots = (int64_t *)malloc(sizeof(int64_t) * slots);
    m->n_slots = slots;
    for (int64_t i = 0; i < slots; ++i) m->slots[i] = -1;
    for (int64_t e = 0; e < w; ++e) {
        uint64_t h = mix((uint64_t)m->keys[e]) & (uint64_t)(slots - 1);
        while (m->slots[h] != -1) h = (h + 1) & (uint64_t)(slots - 1);
        m->slots[h] = e;
    }
    m->n_used = w;
}

static int64_t omap_find(const omap *m, int64_t key) {
    uint64_t mask = (uint64_t)(m->n_slots - 1), h = mix((uint64_t)key) & mask;
    for (;;) {
        int64_t s = m->slots[h];
        if (s == -1) return -1;
        if (s >= 0 && m->keys[s] == key) return s;
        h = (h + 1) & mask;
    }
}

static void omap_set(omap *m, int64_t key, double v) {
    int64_t e = omap_find(m, key);
    if (e >= 0) { m->vals[e] = v; return; } /* Map.set on an existing key keeps its slot */
    if (m->n_entries == m->cap_entries || (m->n_used + 1) * 2 > m->n_slots) omap_rebuild(m, m->n_live + 1);
    e = m->n_entries++;
    m->keys[e] = key;
    m->vals[e] = v;
    m->n_live++;
    uint64_t mask = (uint64_t)(m->n_slots - 1), h = mix((uint64_t)key) & mask;
    while (m->slots[h] >= 0) h = (h + 1) & mask;
    if (m->slots[h] == -1) m->n_used++;
    m->slots[h] = e;
}

static void omap_delete(omap *m, int64_t key) {
    uint64_t mask = (uint64_t)(m->n_slots - 1), h = mix((uint64_t)key) & mask;
    for (;;) {
        int64_t s = m->slots[h];
        if (s == -1) return;
        if (s >= 0 && m->keys[s] == key) {
            m->keys[s] = -1;
            m->slots[h] = -2;
            m->n_live--;
            return;
        }
        h = (h + 1) & mask;
    }
}

/* ------------------------------------------------------------------ store */
typedef struct ostore {
    int64_t size;
    int type;        /* 0 int32, 1 uint32, 2 float32, 3 float64 */
    int default_nan; /* in-memory.js:56-57 */
    omap map;
} ostore;

ostore *ostore_new(int64_t size, int type, int default_nan) { /* in-memory.js:48-64 */
    ostore *s = (ostore *)malloc(sizeof(ostore));
    s->size = size;
    s->type = type;
    s->default_nan = default_nan;
    omap_init(&s->map, 16);
    return s;
}

void ostore_free(ostore *s) {
    if (!s) return;
    omap_free(&s->map);
    free(s);
}

static double default_of(const ostore *s) { return s->default_nan ? NAN : 0.0; }

double ostore_get_value(const ostore *s, int64_t idx) { /* in-memory.js:118-120 */
    int64_t e = omap_find(&s->map, idx);
    return e >= 0 ? s->map.vals[e] : default_of(s);
}

void ostore_set_value(ostore *s, int64_t idx, double v) { /* in-memory.js:122-133 */
    int is_default = s->default_nan ? isnan(v) : (v == 0.0);
    if (!is_default) omap_set(&s->map, idx, v);
    else omap_delete(&s->map, idx);
}

ostore *ostore_clone(const ostore *s) { /* in-memory.js:66-73 */
    ostore *c = ostore_new(s->size, s->type, s->default_nan);
    for (int64_t e = 0; e < s->map.n_entries; ++e)
        if (s->map.keys[e] >= 0) omap_set(&c->map, s->map.keys[e], s->map.vals[e]);
    return c;
}

int ostore_set_data(ostore *s, const double *values, int64_t n) { /* in-memory.js:39-46 */
    if (n != s->size) return -1;
    for (int64_t i = 0; i < n; ++i) ostore_set_value(s, i, values[i]);
    return 0;
}

int ostore_set_data_f32(ostore *s, const float *values, int64_t n) {
    if (n != s->size) return -1;
    for (int64_t i = 0; i < n; ++i) ostore_set_value(s, i, (double)values[i]);
    return 0;
}

void ostore_get_data(const ostore *s, double *out) { /* in-memory.js:30-37 */
    const double d = default_of(s);
    for (int64_t i = 0; i < s->size; ++i) out[i] = d;
    for (int64_t e = 0; e < s->map.n_entries; ++e)
        if (s->map.keys[e] >= 0) out[s->map.keys[e]] = s->map.vals[e];
}

void ostore_fill(ostore *s, double v) { /* in-memory.js:135-137 */
    for (int64_t i = 0; i < s->size; ++i) ostore_set_value(s, i, v);
}

double ostore_total(const ostore *s) { /* in-memory.js:22-28 */
    double t = 0.0;
    for (int64_t e = 0; e < s->map.n_entries; ++e)
        if (s->map.keys[e] >= 0) t += s->map.vals[e];
    return t;
}

int64_t ostore_size(const ostore *s) { return s->size; }
int64_t ostore_count(const ostore *s) { return s->map.n_live; }

/* keys/values in Map (insertion) order */
void ostore_entries(const ostore *s, int64_t *keys, double *vals) {
    int64_t w = 0;
    for (int64_t e = 0; e < s->map.n_entries; ++e)
        if (s->map.keys[e] >= 0) { keys[w] = s->map.keys[e]; vals[w] = s->map.vals[e]; ++w; }
}

static int64_t prod(const int64_t *len, int n) {
    int64_t p = 1;
    for (int i = 0; i < n; ++i) p *= len[i];
    return p;
}

static double js_max(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    if (a == 0.0 && b == 0.0) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static double js_min(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    if (a == 0.0 && b == 0.0) return signbit(a) ? a : b;
    return a < b ? a : b;
}

/* in-memory.js:178-211 */
ostore *ostore_reorder(const ostore *s, int ndim, const int64_t *old_len, const int32_t *new_to_old) {
    ostore *out = ostore_new(s->size, s->type, s->default_nan);
    int64_t coord[64], new_len[64];
    for (int i = 0; i < ndim; ++i) new_len[i] = old_len[new_to_old[i]];
    for (int64_t e = 0; e < s->map.n_entries; ++e) {
        if (s->map.keys[e] < 0) continue;
        int64_t rest = s->map.keys[e];
        for (int i = ndim - 1; i >= 0; --i) { coord[i] = rest % old_len[i]; rest /= old_len[i]; }
        int64_t idx = 0;
        for (int i = 0; i < ndim; ++i) idx = idx * new_len[i] + coord[new_to_old[i]];
        ostore_set_value(out, idx, s->map.vals[e]);
    }
    return out;
}

/* in-memory.js:213-263; keep[d][j] = old item index of new item j */
ostore *ostore_dice(const ostore *s, int ndim, const int64_t *old_len, const int64_t *new_len,
                    const int32_t *const *keep) {
    ostore *out = ostore_new(prod(new_len, ndim), s->type, s->default_nan);
    int64_t *old_to_new[64];
    for (int d = 0; d < ndim; ++d) {
        old_to_new[d] = (int64_t *)malloc(sizeof(int64_t) * (old_len[d] > 0 ? old_len[d] : 1));
        for (int64_t i = 0; i < old_len[d]; ++i) old_to_new[d][i] = -1;
        for (int64_t j = 0; j < new_len[d]; ++j) old_to_new[d][keep[d][j]] = j;
    }
    int64_t coord[64];
    for (int64_t e = 0; e < s->map.n_entries; ++e) {
        if (s->map.keys[e] < 0) continue;
        int64_t rest = s->map.keys[e];
        int halt = 0;
        for (int i = ndim - 1; i >= 0; --i) {
            int64_t nc = old_to_new[i][rest % old_len[i]];
            if (nc < 0) { halt = 1; break; }
            coord[i] = nc;
            rest /= old_len[i];
        }
        if (halt) continue;
        int64_t idx = 0;
        for (int i = 0; i < ndim; ++i) idx = idx * new_len[i] + coord[i];
        ostore_set_value(out, idx, s->map.vals[e]);
    }
    for (int d = 0; d < ndim; ++d) free(old_to_new[d]);
    return out;
}

/* in-memory.js:265-334; maps[d][i] = new item index of old item i */
ostore *ostore_drill_up(const ostore *s, int ndim, const int64_t *old_len, const int64_t *new_len,
                        const int32_t *const *maps, int method) {
    const int64_t new_size = prod(new_len, ndim);
    ostore *out = ostore_new(new_size, s->type, s->default_nan);
    uint16_t *contributions = (uint16_t *)calloc((size_t)(new_size > 0 ? new_size : 1), sizeof(uint16_t)); /* :278 */
    int64_t coord[64];
    for (int64_t e = 0; e < s->map.n_entries; ++e) {
        if (s->map.keys[e] < 0) continue;
        const double old_value = s->map.vals[e];
        int64_t rest = s->map.keys[e];
        for (int i = ndim - 1; i >= 0; --i) { coord[i] = rest % old_len[i]; rest /= old_len[i]; }
        int64_t idx = 0;
        for (int i = 0; i < ndim; ++i) idx = idx * new_len[i] + maps[i][coord[i]];
        int64_t slot = omap_find(&out->map, idx);
        if (slot < 0) {
            ostore_set_value(out, idx, old_value);
        } else {
            const double a = out->map.vals[slot];
            double r;
            switch (method) {
                case O_SUM: case O_AVERAGE: r = a + old_value; break;
                case O_HIGHEST: r = js_max(a, old_value); break;
                case O_LOWEST: r = js_min(a, old_value); break;
                case O_FIRST: r = a; break;
                case O_LAST: r = old_value; break;
                default: r = a * old_value; break;
            }
            ostore_set_value(out, idx, r);
        }
        contributions[idx] += 1;
    }
    if (method == O_AVERAGE) { /* in-memory.js:323-331 */
        for (int64_t idx = 0; idx < new_size; ++idx)
            if (contributions[idx]) ostore_set_value(out, idx, ostore_get_value(out, idx) / contributions[idx]);
    }
    free(contributions);
    return out;
}

/* in-memory.js:336-430; maps[d][j] = old item index of new item j.  `dist` may be NULL;
 * a NaN entry means missing: *err_index receives the index and NULL is returned. */
ostore *ostore_drill_down(const ostore *s, int ndim, const int64_t *old_len, const int64_t *new_len,
                          const int32_t *const *maps, int method_is_sum, const double *dist, int64_t dist_len,
                          int64_t *err_index) {
    const int use_rounding = s->type == 0 || s->type == 1; /* :343 */
    const int64_t old_size = s->size, new_size = prod(new_len, ndim);
    uint32_t *ids = (uint32_t *)calloc((size_t)(old_size > 0 ? old_size : 1), 4);
    uint32_t *total = (uint32_t *)calloc((size_t)(old_size > 0 ? old_size : 1), 4);
    int64_t *idx_new_old = (int64_t *)malloc(sizeof(int64_t) * (size_t)(new_size > 0 ? new_size : 1));
    int64_t coord[64];
    if (err_index) *err_index = -1;
    for (int64_t ni = 0; ni < new_size; ++ni) { /* :361-379 */
        int64_t rest = ni;
        for (int i = ndim - 1; i >= 0; --i) { coord[i] = rest % new_len[i]; rest /= new_len[i]; }
        int64_t oi = 0;
        for (int j = 0; j < ndim; ++j) oi = oi * old_len[j] + maps[j][coord[j]];
        idx_new_old[ni] = oi;
        total[oi] += 1;
    }
    ostore *out = ostore_new(new_size, s->type, s->default_nan);
    for (int64_t ni = 0; ni < new_size; ++ni) { /* :383-427 */
        const int64_t oi = idx_new_old[ni];
        const int64_t slot = omap_find(&s->map, oi);
        if (slot < 0) continue;
        const double old_value = s->map.vals[slot];
        if (old_value == 0.0 || isnan(old_value)) continue; /* `if (!oldValue) continue` */
        const double n = (double)total[oi];
        if (dist) {
            const double added = (double)new_size / (double)old_size;
            const double shared = (double)dist_len / added;
            const int64_t di = (int64_t)(floor((double)ni / ((double)new_size / shared)) * added + fmod((double)ni, added));
            if (di < 0 || di >= dist_len || isnan(dist[di])) {
                if (err_index) *err_index = di;
                ostore_free(out);
                out = NULL;
                break;
            }
            ostore_set_value(out, ni, old_value * dist[di]);
        } else if (method_is_sum) {
            if (use_rounding) {
                const double value = floor(old_value / n);
                const double remainder = fmod(old_value, n);
                const double k = (double)ids[oi];
                const double step = remainder / n;
                const int last_is_same = floor(k * step) == floor((k - 1.0) * step);
                ostore_set_value(out, ni, last_is_same ? value : value + 1.0);
            } else {
                ostore_set_value(out, ni, old_value / n);
            }
        } else {
            ostore_set_value(out, ni, old_value);
        }
        ids[oi]++;
    }
    free(ids); free(total); free(idx_new_old);
    return out;
}

/* in-memory.js:139-176; his_to_mine[d][j] = my item index of his item j, -1 when unknown
 * (the reference then indexes with NaN and the write never lands in a real cell). */
void ostore_load(ostore *dst, const ostore *src, int ndim, const int64_t *my_len, const int64_t *his_len,
                 const int32_t *const *his_to_mine) {
    int64_t coord[64];
    for (int64_t hi = 0; hi < src->size; ++hi) {
        int64_t rest = hi;
        for (int i = ndim - 1; i >= 0; --i) { coord[i] = rest % his_len[i]; rest /= his_len[i]; }
        int64_t mi = 0;
        int drop = 0;
        for (int i = 0; i < ndim; ++i) {
            const int32_t off = his_to_mine[i][coord[i]];
            if (off < 0) { drop = 1; break; }
            mi = mi * my_len[i] + off;
        }
        if (!drop) ostore_set_value(dst, mi, ostore_get_value(src, hi));
    }
}
