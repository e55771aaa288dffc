/* CPU oracle in plain C for the hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Same contract as oracle/oracle_np.py (read its header): a restatement of the
 * per-row arithmetic of /root/reference/src/deal_yolo_data/core/processor.py on
 * the CSR buffers of DESIGN.md §3, used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker / timed CPU port.  Never linked or
 * loaded by the product path.  Parity is pinned through oracle_np.py (tests hold
 * both to the reference-generated fixtures in tests/golden/).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off keeps every multiply and add separately rounded, like
 * CPython's float arithmetic.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- K1: get_bbox_points, processor.py:252-260 -------------------------- */
/* builtin min/max are left folds that replace the running value only on a
 * strict comparison, so the first of equal values (and of 0.0 / -0.0) survives
 * and a NaN survives only from the first position.                          */
void orc_bbox(const int64_t* poly_off, const double* xy, int64_t n_poly,
              double* pts, uint8_t* valid, int32_t* arg /* may be NULL */) {
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n_poly; ++p) {
        int64_t a = poly_off[p], b = poly_off[p + 1];
        if (b <= a) {
            valid[p] = 0;
            pts[4 * p] = pts[4 * p + 1] = pts[4 * p + 2] = pts[4 * p + 3] = 0.0;
            if (arg) arg[4 * p] = arg[4 * p + 1] = arg[4 * p + 2] = arg[4 * p + 3] = -1;
            continue;
        }
        double mnx = xy[2 * a], mxx = mnx, mny = xy[2 * a + 1], mxy = mny;
        int32_t imnx = 0, imxx = 0, imny = 0, imxy = 0;
        for (int64_t v = a + 1; v < b; ++v) {
            double x = xy[2 * v], y = xy[2 * v + 1];
            int32_t k = (int32_t)(v - a);
            if (x < mnx) { mnx = x; imnx = k; }
            if (x > mxx) { mxx = x; imxx = k; }
            if (y < mny) { mny = y; imny = k; }
            if (y > mxy) { mxy = y; imxy = k; }
        }
        valid[p] = 1;
        pts[4 * p] = mnx; pts[4 * p + 1] = mny; pts[4 * p + 2] = mxx; pts[4 * p + 3] = mxy;
        if (arg) { arg[4 * p] = imnx; arg[4 * p + 1] = imny; arg[4 * p + 2] = imxx; arg[4 * p + 3] = imxy; }
    }
}

/* ---- K2: calculate_iou / extract_boxes / meet_conditions, :328-376 ------- */
static inline double pymin(double a, double b) { return b < a ? b : a; }
static inline double pymax(double a, double b) { return b > a ? b : a; }

static double iou_pair(const double* b1, const double* b2) {
    double xi1 = pymax(b1[0], b2[0]), yi1 = pymax(b1[1], b2[1]);
    double xi2 = pymin(b1[2], b2[2]), yi2 = pymin(b1[3], b2[3]);
    double dx = xi2 - xi1, dy = yi2 - yi1;
    double w = dx > 0 ? dx : 0.0, h = dy > 0 ? dy : 0.0;   /* max(0, d) */
    double inter = w * h;
    if (inter == 0) return 0.0;
    double area1 = (b1[2] - b1[0]) * (b1[3] - b1[1]);
    double area2 = (b2[2] - b2[0]) * (b2[3] - b2[1]);
    double uni = area1 + area2 - inter;
    return uni != 0 ? inter / uni : 0.0;
}

void orc_iou_filter(const int64_t* img_off, const double* pts, const uint8_t* valid,
                    int64_t n_img, int64_t min_boxes, double thr,
                    uint8_t* high, int32_t* count) {
#pragma omp parallel
    {
        int64_t cap = 64;
        double* bx = (double*)malloc(sizeof(double) * 4 * cap);
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < n_img; ++i) {
            int64_t a = img_off[i], b = img_off[i + 1], n = 0;
            if (b - a > cap) { cap = b - a; bx = (double*)realloc(bx, sizeof(double) * 4 * cap); }
            for (int64_t q = a; q < b; ++q) {          /* prefix up to the first null bbox */
                if (valid && !valid[q]) break;
                const double* p = pts + 4 * q;
                bx[4 * n] = pymin(p[0], p[2]); bx[4 * n + 1] = pymin(p[1], p[3]);
                bx[4 * n + 2] = pymax(p[0], p[2]); bx[4 * n + 3] = pymax(p[1], p[3]);
                ++n;
            }
            count[i] = (int32_t)n;
            uint8_t hit = 0;
            if (n >= min_boxes) {
                for (int64_t s = 0; s < n && !hit; ++s)
                    for (int64_t t = s + 1; t < n; ++t)
                        if (iou_pair(bx + 4 * s, bx + 4 * t) >= thr) { hit = 1; break; }
            }
            high[i] = hit;
        }
        free(bx);
    }
}

/* ---- K0: this repo's 64-bit string hash --------------- */
#define HM 0xC6A4A7935BD1E995ULL
#define HSEED 0x8445D61A4E774912ULL

static uint64_t hash_bytes(const uint8_t* b, int64_t n) {
    uint64_t h = HSEED ^ ((uint64_t)n * HM);
    int64_t nblk = n / 8;
    for (int64_t i = 0; i < nblk; ++i) {
        uint64_t k; memcpy(&k, b + 8 * i, 8);
        k *= HM; k ^= k >> 47; k *= HM;
        h ^= k; h *= HM;
    }
    int64_t rem = n - 8 * nblk;
    if (rem) {
        uint64_t t = 0; memcpy(&t, b + 8 * nblk, (size_t)rem);
        h ^= t; h *= HM;
    }
    h ^= h >> 47; h *= HM; h ^= h >> 47;
    return h;
}

void orc_hash_strings(const int64_t* off, const uint8_t* bytes, int64_t n, uint64_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) out[r] = hash_bytes(bytes + off[r], off[r + 1] - off[r]);
}

/* ---- open-addressing table shared by K4/K5 (host, sequential) ----------- */
typedef struct { uint64_t* key; int64_t* lo; int64_t* hi; int64_t* cnt; uint8_t* used; uint64_t mask; } tab_t;

static int tab_init(tab_t* t, int64_t n) {
    uint64_t cap = 16; while (cap < (uint64_t)n * 2) cap <<= 1;
    t->mask = cap - 1;
    t->key = (uint64_t*)malloc(cap * 8); t->lo = (int64_t*)malloc(cap * 8);
    t->hi = (int64_t*)malloc(cap * 8); t->cnt = (int64_t*)calloc(cap, 8);
    t->used = (uint8_t*)calloc(cap, 1);
    return t->key && t->lo && t->hi && t->cnt && t->used ? 0 : -1;
}
static void tab_free(tab_t* t) { free(t->key); free(t->lo); free(t->hi); free(t->cnt); free(t->used); }
static uint64_t tab_slot(const tab_t* t, uint64_t k, int insert) {
    uint64_t s = (k * 0x9E3779B97F4A7C15ULL) >> 20 & t->mask;
    for (;;) {
        if (!t->used[s]) return insert ? s : UINT64_MAX;
        if (t->key[s] == k) return s;
        s = (s + 1) & t->mask;
    }
}

/* ---- K4: drop_duplicates(subset=["source"], keep=...), :140-144 --------- */
/* keep_mode: 0 first, 1 last, 2 False (drop every member of a repeated group).
 * Null (NaN) cells form one group.                                            */
int orc_dedup(const uint64_t* keys, const uint8_t* null, int64_t n, int keep_mode,
              uint8_t* keep, int64_t* rep) {
    tab_t t; if (tab_init(&t, n) != 0) return -1;
    int64_t nlo = -1, nhi = -1, ncnt = 0;
    for (int64_t r = 0; r < n; ++r) {
        if (null && null[r]) { if (nlo < 0) nlo = r; nhi = r; ++ncnt; continue; }
        uint64_t s = tab_slot(&t, keys[r], 1);
        if (!t.used[s]) { t.used[s] = 1; t.key[s] = keys[r]; t.lo[s] = r; }
        t.hi[s] = r; t.cnt[s]++;
    }
    for (int64_t r = 0; r < n; ++r) {
        int64_t lo, hi, c;
        if (null && null[r]) { lo = nlo; hi = nhi; c = ncnt; }
        else { uint64_t s = tab_slot(&t, keys[r], 0); lo = t.lo[s]; hi = t.hi[s]; c = t.cnt[s]; }
        if (keep_mode == 0) { rep[r] = lo; keep[r] = lo == r; }
        else if (keep_mode == 1) { rep[r] = hi; keep[r] = hi == r; }
        else { rep[r] = lo; keep[r] = c == 1; }
    }
    tab_free(&t);
    return 0;
}

/* ---- K5: ~main.isin(set(ref.dropna())), :194-199 ------------------------ */
int orc_antijoin(const uint64_t* mk, const uint8_t* mnull, int64_t n_main,
                 const uint64_t* rk, const uint8_t* rnull, int64_t n_ref,
                 uint8_t* keep, int64_t* ref_row) {
    tab_t t; if (tab_init(&t, n_ref) != 0) return -1;
    for (int64_t r = 0; r < n_ref; ++r) {
        if (rnull && rnull[r]) continue;
        uint64_t s = tab_slot(&t, rk[r], 1);
        if (!t.used[s]) { t.used[s] = 1; t.key[s] = rk[r]; t.lo[s] = r; }
    }
    for (int64_t r = 0; r < n_main; ++r) {
        keep[r] = 1; ref_row[r] = -1;
        if (mnull && mnull[r]) continue;
        uint64_t s = tab_slot(&t, mk[r], 0);
        if (s != UINT64_MAX) { keep[r] = 0; ref_row[r] = t.lo[s]; }
    }
    tab_free(&t);
    return 0;
}

/* ---- K3: name rewrite via LUT, processor.py:582-602 ---------------------- */
/* counters: [0] total_objects [1] missing_name [2] total_labels
 *           [3] replaced_labels [4] replaced_objects [5] replaced_rows       */
void orc_label_lut(const int64_t* img_off, const int32_t* label_id, int64_t n_img,
                   const int32_t* lut_new, const int32_t* lut_ntok, const int32_t* lut_nrep,
                   int32_t* new_id, uint8_t* row_rep, uint64_t* counters) {
    uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
#pragma omp parallel for schedule(static) reduction(+ : c0, c1, c2, c3, c4, c5)
    for (int64_t i = 0; i < n_img; ++i) {
        uint8_t rr = 0;
        for (int64_t q = img_off[i]; q < img_off[i + 1]; ++q) {
            int32_t v = label_id[q];
            ++c0; new_id[q] = v;
            if (v < 0) { ++c1; continue; }
            c2 += (uint64_t)lut_ntok[v];
            if (lut_nrep[v] > 0) { new_id[q] = lut_new[v]; c3 += (uint64_t)lut_nrep[v]; ++c4; rr = 1; }
        }
        row_rep[i] = rr; c5 += rr;
    }
    counters[0] = c0; counters[1] = c1; counters[2] = c2; counters[3] = c3; counters[4] = c4; counters[5] = c5;
}

/* ---- K6: category expansion, processor.py:751-775 ------------------------ */
/* Two passes (count, then stable fill); cat_off has n_cat+1 entries.         */
void orc_split_expand(const int64_t* img_off, const int32_t* label_id, int64_t n_img,
                      const int32_t* cat_of_label, int32_t n_cat,
                      int64_t* cat_off, int64_t* exp_img, int64_t* exp_box, int32_t* exp_cat) {
    int64_t* cur = (int64_t*)calloc((size_t)n_cat + 1, 8);
    for (int64_t i = 0; i < n_img; ++i)
        for (int64_t q = img_off[i]; q < img_off[i + 1]; ++q) {
            int32_t v = label_id[q]; if (v < 0) continue;
            int32_t c = cat_of_label[v]; if (c >= 0) cur[c + 1]++;
        }
    cat_off[0] = 0;
    for (int32_t c = 0; c < n_cat; ++c) cat_off[c + 1] = cat_off[c] + cur[c + 1];
    if (exp_img) {
        for (int32_t c = 0; c < n_cat; ++c) cur[c] = cat_off[c];
        for (int64_t i = 0; i < n_img; ++i)
            for (int64_t q = img_off[i]; q < img_off[i + 1]; ++q) {
                int32_t v = label_id[q]; if (v < 0) continue;
                int32_t c = cat_of_label[v]; if (c < 0) continue;
                int64_t d = cur[c]++;
                exp_img[d] = i; exp_box[d] = q; exp_cat[d] = c;
            }
    }
    free(cur);
}
