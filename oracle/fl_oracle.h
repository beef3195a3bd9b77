/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * fl_oracle: an OpenCV-free, CPU-only C restatement of the FEALESS LINE-MOD + ICP hot path
 * (reference rlvc/FEALESS, files linemod/linemod.cpp, ICP/ sources).  It is the checker the parity tests
 * compare the CUDA path with, and the "port" CPU baseline bench.py times.  Nothing under
 * fealess_b200/ may include, link or call it.
 *
 * PARITY STATUS: pinned on the reference itself.  oracle/_ref/libfl_ref.so is the reference's own linemod.cpp and ICP
 * sources compiled unmodified (oracle/build_ref.py; OpenCV replaced by the stand-in oracle/ref_shim, whose primitives are
 * pinned on the real cv2 4.13 and can be swapped for it at run time); tests/test_oracle_ref.py shows this file equal to it
 * bit for bit on every integer stage, on the pre-sort match lists, on depthTo3d / matToVec / NMS and on the ICP poses.
 * Older anchors kept: the sha256 of the reference's two embedded tables and the fixtures in tests/golden/ written by
 * oracle/oracle_cv2.py (the cited lines evaluated on cv2).  The reference holds no tests or golden vectors of its own.
 *
 * All file:line citations are relative to /root/reference.
 */
#ifndef FL_ORACLE_H
#define FL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t x, y; float similarity; int32_t class_idx, template_id; } flo_match_t;
typedef struct flo_detector flo_detector;

/* ---- embedded tables (closed forms; pinned by sha256) ---- */
void flo_normal_lut(uint8_t out[8000]);          /* linemod/normal_lut.i:4 */
void flo_similarity_lut(uint8_t out[256]);       /* linemod/linemod.cpp:970 */

/* ---- stage functions ---- */
void flo_gaussian7_bgr(const uint8_t* src, int W, int H, uint8_t* dst);                  /* cv::GaussianBlur 7x7 s=0, linemod.cpp:247 */
void flo_sobel3_bgr(const uint8_t* src, int W, int H, int16_t* dx, int16_t* dy);         /* cv::Sobel k3, :248-249 */
void flo_phase_q16(const float* dx, const float* dy, int n, uint8_t* q);                 /* cv::phase + convertTo(16/360), :303, :314 */
void flo_color_quantize(const uint8_t* bgr, int W, int H, float weak_thr, uint8_t* q, float* mag_or_null); /* :230-385 */
void flo_pyrdown_bgr(const uint8_t* src, int W, int H, uint8_t* dst);                    /* cv::pyrDown, :443 */
void flo_resize_nn_half_u8(const uint8_t* src, int W, int H, uint8_t* dst);              /* resize NEAREST, :448, 731, 736 */
void flo_median5_u8(const uint8_t* src, int W, int H, uint8_t* dst);                     /* cv::medianBlur 5, :684 */
void flo_depth_quantize(const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* out); /* :595-685 */
void flo_spread(const uint8_t* q, int W, int H, int T, uint8_t* out);                    /* :950-965 */
void flo_response_maps(const uint8_t* sp, int n, uint8_t* out8);                         /* :979-1048; out8 = [8][n] */
void flo_linearize(const uint8_t* resp, int W, int H, int T, uint8_t* out);              /* :1060-1088 */

/* ---- detector (Detector::match, linemod.cpp:1356-1577) ---- */
/* modality_kind: 0 = ColorGradient (source = BGR u8), 1 = DepthNormal (source = depth u16) */
flo_detector* flo_detector_create(int n_levels, const int* T, int n_modalities, const int* modality_kind,
                                  float weak_threshold, int distance_threshold, int difference_threshold);
void flo_detector_destroy(flo_detector* d);
/* headers: 7 ints per (template, level, modality): width,height,offset_x,offset_y,pyramid_level,feature_begin,feature_count
 * features: 3 ints each (x,y,label); class_of: class index per template (class-contiguous). Returns 0, or <0 on bad input. */
int flo_detector_set_templates(flo_detector* d, int n_templates, const int32_t* headers, const int32_t* features,
                               int n_features, const int32_t* class_of);
/* front end for one frame: quantise, spread, response, linearise for every level/modality. masks: NULL or
 * n_modalities pointers (each NULL or W*H u8).  Returns 0; -2 if a level violates W%T / H%T / (W*H)%16 (:981, :1062). */
int flo_detector_process(flo_detector* d, const uint8_t* bgr, const uint16_t* depth, int W, int H,
                         const uint8_t* const* masks);
const uint8_t* flo_detector_quantized(const flo_detector* d, int level, int modality, int* W, int* H);
const uint8_t* flo_detector_spread(const flo_detector* d, int level, int modality);
/* linear memory of (level, modality, label): T*T rows of cols bytes, continuous */
const uint8_t* flo_detector_lm(const flo_detector* d, int level, int modality, int label, int* rows, int* cols);
/* global similarity of one template at the coarsest level, summed over modalities (u16 [H'*W']) - debug/parity */
int flo_detector_similarity(const flo_detector* d, int template_idx, uint16_t* out);
/* matchClass over all (or class_filter'ed) templates on the processed frame.
 * canonical != 0: sort by (similarity desc, template_id asc, class asc, y asc, x asc) + adjacent unique on
 * (x,y,similarity,class); canonical == 0: raw emission order (template-major, row-major cells).
 * n_threads > 1 parallelises over templates (NOT the reference's execution model; result identical).
 * Returns the number of matches written (<= cap); *n_total gets the number before truncation. */
int flo_detector_match_templates(const flo_detector* d, float threshold, const int32_t* class_filter, int n_filter,
                                 int canonical, int n_threads, flo_match_t* out, int cap, int* n_total);

/* ---- template training: Detector::addTemplate (linemod.cpp:1579-1615; extractTemplate :461-513, :747-825; selectScatteredFeatures
 * :134-163; cropTemplates :52-96) ---- */
void flo_erode3_u8(const uint8_t* src, int W, int H, int iterations, uint8_t* dst);      /* cv::erode 3x3 BORDER_REPLICATE, :466, 753 */
void flo_distance_c3(const uint8_t* src, int W, int H, float* dst);                      /* cv::distanceTransform(DIST_C, 3), :765 */
/* one view -> one template pyramid (headers / features in the layout of flo_detector_set_templates, one template); num_features per
 * modality at level 0 (default 63), strong_threshold 55, extract_threshold 2.  Returns 0, or -1 when a level has too few candidates. */
int flo_add_template(const flo_detector* d, const uint8_t* bgr, const uint16_t* depth, const uint8_t* mask_or_null, int W, int H,
                     const int* num_features_per_modality, float strong_threshold, int extract_threshold,
                     int32_t* headers, int32_t* features, int feature_cap, int* n_features, int32_t bbox[4]);

/* ---- ICP (ICP/depth_to_3d.cpp, common.cpp, ICP.cpp, detection.cpp, NMS.cpp) ---- */
void flo_depth_to_3d_mm(const uint16_t* depth, int W, int H, float fx, float fy, float cx, float cy, float* out3);
int flo_pair_points(const float* ref3, const float* mod3, int W, int H, const int rect_ref[4], const int rect_mod[4],
                    float* pts_ref, float* pts_mod);
/* returns dist_mean (or -1 if < 3 points). trace (nullable): up to 3*it_thr floats (dist_mean, dist_diff, n_cor per iteration) */
float flo_icp_cloud_to_cloud_ex(const float* pts_ref, int n_ref, const float* pts_model, int n_model,
                                float R[9], float T[3], float* inlier_ratio, int icp_it_thr, float dist_mean_thr,
                                float dist_diff_thr, int* iterations, float* trace);
/* detection(): K_ref = fx,fy,cx,cy ; rect = x,y,w,h.  Returns 0, -3 if a rect leaves the image (cv ROI throw, detection.cpp:43-44). */
int flo_detection(const uint16_t* model_depth, const uint16_t* ref_depth, int W, int H, const float K_ref[4],
                  const int rect_model[4], const int rect_ref[4], int icp_it_thr, float dist_mean_thr,
                  float dist_diff_thr, const float r_match[9], const float t_match[3], float d_match,
                  float T_final[3], float R_final[9], float* dist_mean, float* inlier_ratio, int* iterations,
                  int* n_points);
/* nonMaximumSuppression (NMS.cpp:6-39): returns the count; out_idx = index of the emitted object per group */
int flo_nms(const float* t3, const int32_t* n_model_pts, const float* icp_dist, int n, float th_obj_dist, int32_t* out_idx);
void flo_svd3_rot(const float cov[9], float R[9]);   /* R = V * U^T of cv::SVD::compute(cov), ICP.cpp:741-744 */

#ifdef __cplusplus
}
#endif
#endif
