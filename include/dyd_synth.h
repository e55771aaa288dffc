/* dyd_synth.h -- C ABI of libdyd_synth.so: the seeded synthetic-table generator used by bench.py and
 * the at-scale GPU tests.  Test / measurement infrastructure, deliberately NOT part of
 * the product library (libdyd.so, include/dyd.h): the reference has no generator.                     */
#ifndef DYD_SYNTH_H_
#define DYD_SYNTH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------ synthetic tables (§8d) ---
 * Device-side twin of deal_yolo_daya_b200/synth.py (bit-identical output).       */
int dyd_synth_counts(uint64_t seed, int64_t first_img, int64_t n_img, const uint64_t* d_pois_thr,
                     int32_t n_thr, int64_t* d_npoly /* [n_img] */, void* stream);
int dyd_synth_nvert(uint64_t seed, int64_t first_img, int64_t n_img, const int64_t* d_img_off,
                    int64_t* d_nvert /* [n_poly] */, void* stream);
int dyd_synth_fill(uint64_t seed, int64_t first_img, int64_t n_img, const int64_t* d_img_off,
                   const int64_t* d_poly_off, double* d_xy, int32_t* d_label_id, void* stream);
int dyd_synth_urls(uint64_t seed, int64_t first_row, int64_t n, int64_t n_main_for_ref /* <0: main table */,
                   int64_t* d_url_id, int64_t* d_len /* [n] byte length of each URL */, void* stream);
int dyd_synth_url_bytes(const int64_t* d_url_id, const int64_t* d_off, int64_t n, uint8_t* d_bytes, void* stream);
int dyd_synth_crowd(uint64_t seed, int64_t first_img, int64_t n_img, int32_t lo, int32_t hi,
                    const int64_t* d_img_off /* NULL: write counts to d_nbox */, int64_t* d_nbox,
                    double* d_pts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYD_SYNTH_H_ */
