// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// C entry points (flr_*) into the REFERENCE's own ICP / NMS / back-projection code: /root/reference/ICP/{ICP,NMS,common,
// detection,depth_to_3d}.cpp are compiled unmodified from where they lie (oracle/build_ref.py) and linked with this file.
// The functions called here are the reference's public ones (ICP/ICP.h, ICP/NMS.h, ICP/detection.h, ICP/depth_to_3d.h,
// ICP/common.h).
#include <cstdint>

#include "ICP.h"
#include "NMS.h"
#include "common.h"
#include "depth_to_3d.h"
#include "detection.h"

namespace {
struct Quiet {   // the reference prints timing lines with std::cout on every call (detection.cpp:145, 223, 252)
  std::ios_base::iostate saved;
  Quiet() : saved(std::cout.rdstate()) { std::cout.setstate(std::ios_base::failbit); }
  ~Quiet() { std::cout.clear(saved); }
};
}  // namespace

extern "C" {

void flr_set_hook(int op, cvshim::hook_fn f) { cvshim::set_hook(op, f); }

// cup_d2pc::depthTo3d (depth_to_3d.cpp:190-221, 244-269) then scale_mat_vec3f(.., 1000) (common.cpp:382-389), as
// detection() does (detection.cpp:28-40).  K = (fx, fy, cx, cy) as floats.
int flr_depth_to_3d_mm(const uint16_t* depth, int W, int H, float fx, float fy, float cx, float cy, float* out3) {
  try {
    cv::Mat img(H, W, CV_16UC1, (void*)depth);
    cv::Mat_<float> K(3, 3, 0.f);
    K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy; K(2, 2) = 1.f;
    cv::Mat_<cv::Vec3f> pts;
    cup_d2pc::depthTo3d(img, K, pts);
    scale_mat_vec3f(pts, 1000);
    for (int r = 0; r < H; ++r) std::memcpy(out3 + (size_t)r * W * 3, pts.ptr(r), sizeof(float) * 3 * W);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}

// matToVec(ref ROI, model ROI, ..) (common.cpp:395-416) on two W x H x 3 clouds; rect = x, y, w, h.  Returns the pair count,
// -3 if a rect leaves the image (cv ROI assert, detection.cpp:43-44).
int flr_pair_points(const float* ref3, const float* mod3, int W, int H, const int rect_ref[4], const int rect_mod[4], float* pts_ref,
                    float* pts_mod) {
  try {
    cv::Mat_<cv::Vec3f> R(cv::Mat(H, W, CV_32FC3, (void*)ref3)), M(cv::Mat(H, W, CV_32FC3, (void*)mod3));
    cv::Mat_<cv::Vec3f> r = R(cv::Rect(rect_ref[0], rect_ref[1], rect_ref[2], rect_ref[3]));
    cv::Mat_<cv::Vec3f> m = M(cv::Rect(rect_mod[0], rect_mod[1], rect_mod[2], rect_mod[3]));
    std::vector<cv::Vec3f> pr, pm;
    matToVec(r, m, pr, pm);
    for (size_t i = 0; i < pr.size(); ++i) for (int k = 0; k < 3; ++k) { pts_ref[3 * i + k] = pr[i][k]; pts_mod[3 * i + k] = pm[i][k]; }
    return (int)pr.size();
  } catch (const cv::Exception&) { return -3; }
}

// icpCloudToCloud_Ex (ICP.cpp:617-809)
float flr_icp_cloud_to_cloud_ex(const float* pts_ref, int n_ref, const float* pts_model, int n_model, float R[9], float T[3],
                                float* inlier_ratio, int icp_it_thr, float dist_mean_thr, float dist_diff_thr) {
  Quiet q;
  std::vector<cv::Vec3f> pr(n_ref), pm(n_model);
  for (int i = 0; i < n_ref; ++i) pr[i] = cv::Vec3f(pts_ref[3 * i], pts_ref[3 * i + 1], pts_ref[3 * i + 2]);
  for (int i = 0; i < n_model; ++i) pm[i] = cv::Vec3f(pts_model[3 * i], pts_model[3 * i + 1], pts_model[3 * i + 2]);
  cv::Matx33f Rm; cv::Vec3f Tv; float ratio = 0.f;
  float dm = icpCloudToCloud_Ex(pr, pm, Rm, Tv, ratio, icp_it_thr, dist_mean_thr, dist_diff_thr);
  for (int k = 0; k < 9; ++k) R[k] = Rm.val[k];
  for (int k = 0; k < 3; ++k) T[k] = Tv[k];
  if (inlier_ratio) *inlier_ratio = ratio;
  return dm;
}

// detection() (detection.cpp:11-254).  K_ref = fx, fy, cx, cy; rect = x, y, w, h.  Returns 0, -3 when OpenCV would throw
// (a rect outside the frame, detection.cpp:43-44).
int flr_detection(const uint16_t* model_depth, const uint16_t* ref_depth, int W, int H, const float K_ref[4], const int rect_model[4],
                  const int rect_ref[4], int icp_it_thr, float dist_mean_thr, float dist_diff_thr, const float r_match[9],
                  const float t_match[3], float d_match, float T_final[3], float R_final[9]) {
  try {
    Quiet q;
    cv::Mat model(H, W, CV_16UC1, (void*)model_depth), ref(H, W, CV_16UC1, (void*)ref_depth);
    TCamIntrinsicParam cam;
    cam.nWidth = W; cam.nHeight = H; cam.dFx = K_ref[0]; cam.dFy = K_ref[1]; cam.dCx = K_ref[2]; cam.dCy = K_ref[3];
    cv::Matx33f rm; for (int k = 0; k < 9; ++k) rm.val[k] = r_match[k];
    cv::Vec3f tm(t_match[0], t_match[1], t_match[2]), Tf; cv::Matx33f Rf;
    detection(model, ref, cam, cv::Rect(rect_model[0], rect_model[1], rect_model[2], rect_model[3]),
              cv::Rect(rect_ref[0], rect_ref[1], rect_ref[2], rect_ref[3]), icp_it_thr, dist_mean_thr, dist_diff_thr, rm, tm, d_match, Tf, Rf);
    for (int k = 0; k < 9; ++k) R_final[k] = Rf.val[k];
    for (int k = 0; k < 3; ++k) T_final[k] = Tf[k];
    return 0;
  } catch (const cv::Exception&) { return -3; }
}

// nonMaximumSuppression (NMS.cpp:6-39): out_idx = index of the object each emitted pose came from
int flr_nms(const float* t3, const int32_t* n_model_pts, const float* icp_dist, int n, float th_obj_dist, int32_t* out_idx) {
  std::vector<obj_data> objs(n);
  for (int i = 0; i < n; ++i) {
    objs[i].match_class = i;   // carried through PoseResult::object_id so the emitted pose can be traced back
    objs[i].match_sim = 0.f;
    objs[i].r = cv::Mat(cv::Matx33f::eye());
    objs[i].t = cv::Mat(cv::Vec3f(t3[3 * i], t3[3 * i + 1], t3[3 * i + 2]));
    objs[i].pts_model.resize((size_t)std::max(n_model_pts[i], 0));
    objs[i].icp_dist = icp_dist[i];
    objs[i].check_done = false;
  }
  std::vector<PoseResult> res;
  nonMaximumSuppression(objs, th_obj_dist, res);
  for (size_t i = 0; i < res.size(); ++i) out_idx[i] = res[i].object_id();
  return (int)res.size();
}

// cv::SVD::compute as the ICP calls it (ICP.cpp:741-744): returns R = Mat(vt.t() * u.t())
void flr_svd3_rot(const float cov[9], float R[9]) {
  cv::Matx33f c; for (int k = 0; k < 9; ++k) c.val[k] = cov[k];
  cv::Mat w, u, vt;
  cv::SVD::compute(c, w, u, vt);
  cv::Matx33f r = cv::Mat(vt.t() * u.t());
  for (int k = 0; k < 9; ++k) R[k] = r.val[k];
}

// the shim's own primitives, exported so that tests can pin each one on the real cv2 (tests/test_oracle_ref.py)
void flr_prim_gaussian7(const uint8_t* bgr, int W, int H, uint8_t* out) {
  cv::Mat s(H, W, CV_8UC3, (void*)bgr), d; cv::GaussianBlur(s, d, cv::Size(7, 7), 0, 0, cv::BORDER_REPLICATE);
  for (int r = 0; r < H; ++r) std::memcpy(out + (size_t)r * W * 3, d.ptr(r), (size_t)W * 3);
}
void flr_prim_sobel(const uint8_t* bgr, int W, int H, int16_t* dx, int16_t* dy) {
  cv::Mat s(H, W, CV_8UC3, (void*)bgr), a, b;
  cv::Sobel(s, a, CV_16S, 1, 0, 3, 1.0, 0.0, cv::BORDER_REPLICATE); cv::Sobel(s, b, CV_16S, 0, 1, 3, 1.0, 0.0, cv::BORDER_REPLICATE);
  for (int r = 0; r < H; ++r) { std::memcpy(dx + (size_t)r * W * 3, a.ptr(r), (size_t)W * 6); std::memcpy(dy + (size_t)r * W * 3, b.ptr(r), (size_t)W * 6); }
}
void flr_prim_phase_q(const float* dx, const float* dy, int n, float* angle, uint8_t* q) {
  cv::Mat x(1, n, CV_32F, (void*)dx), y(1, n, CV_32F, (void*)dy), a, qq;
  cv::phase(x, y, a, true); a.convertTo(qq, CV_8U, 16.0 / 360.0);
  std::memcpy(angle, a.data, sizeof(float) * n); std::memcpy(q, qq.data, n);
}
void flr_prim_median5(const uint8_t* src, int W, int H, uint8_t* out) {
  cv::Mat s(H, W, CV_8UC1, (void*)src), d; cv::medianBlur(s, d, 5);
  for (int r = 0; r < H; ++r) std::memcpy(out + (size_t)r * W, d.ptr(r), W);
}
void flr_prim_pyrdown(const uint8_t* bgr, int W, int H, uint8_t* out) {
  cv::Mat s(H, W, CV_8UC3, (void*)bgr), d; cv::pyrDown(s, d, cv::Size(W / 2, H / 2));
  for (int r = 0; r < H / 2; ++r) std::memcpy(out + (size_t)r * (W / 2) * 3, d.ptr(r), (size_t)(W / 2) * 3);
}
void flr_prim_resize_nn_half(const uint8_t* src, int W, int H, uint8_t* out) {
  cv::Mat s(H, W, CV_8UC1, (void*)src), d; cv::resize(s, d, cv::Size(W / 2, H / 2), 0.0, 0.0, cv::INTER_NEAREST);
  for (int r = 0; r < H / 2; ++r) std::memcpy(out + (size_t)r * (W / 2), d.ptr(r), W / 2);
}
void flr_prim_knn1(const float* ref, int n_ref, const float* q, int nq, int32_t* idx, float* dist) {
  cvflann::Matrix<float> data((float*)ref, n_ref, 3), qq((float*)q, nq, 3);
  cvflann::Index<cvflann::L2_Simple<float> > index(data, cvflann::KDTreeSingleIndexParams(15));
  index.buildIndex();
  cvflann::Matrix<int> ii(idx, nq, 1); cvflann::Matrix<float> dd(dist, nq, 1);
  index.knnSearch(qq, ii, dd, 1, cvflann::SearchParams());
}

}  // extern "C"
