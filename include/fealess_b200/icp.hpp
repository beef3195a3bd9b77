// fealess_b200/icp.hpp - C++ host-side mirror of the reference's ICP / pose-refinement interface (reference headers:
// ICP/detection.h:9-11, ICP/ICP.h:165-172, ICP/depth_to_3d.h:12-13, ICP/NMS.h:14-16, ICP/obj_data.h:7-17,
// ICP/pose_result.h:7-106, CadReco/lotus_common.h:41-50).  Same free-function names, argument order and units
// (millimetres); all arithmetic happens on the GPU behind the C ABI (include/fealess_b200.h).  Header-only.
#ifndef FEALESS_B200_ICP_HPP
#define FEALESS_B200_ICP_HPP

#include <vector>

#include "../fealess_b200.h"
#include "cv_min.hpp"
#include "linemod.hpp"   // fealess_b200::check_status, device_ordinal

#ifndef CV_64F
#define CV_64F 6
#endif

// CadReco/lotus_common.h:41-50 (guarded so that a FEALESS build keeps its own definition)
#ifndef FEALESS_B200_HAVE_LOTUS_COMMON
struct TCamIntrinsicParam {
  int nWidth, nHeight;
  double dFx, dFy, dCx, dCy;
  std::vector<double> vdDistCoeff;
};
#endif

namespace fealess_b200 {
// process-wide handle for the free functions below (ICP needs no template database); created on first use
inline fl_handle* icp_handle() {
  static fl_handle* h = nullptr;
  if (!h) {
    fl_params_t p;
    fl_default_params(&p);
    p.device = device_ordinal();
    check_status(fl_create(&p, &h), "fl_create");
  }
  return h;
}
inline fl_intrinsics_t intrinsics_of(const TCamIntrinsicParam& c) { fl_intrinsics_t k = {(float)c.dFx, (float)c.dFy, (float)c.dCx, (float)c.dCy}; return k; }
}  // namespace fealess_b200

// detection() (ICP/detection.cpp:11-254): refines one LINE-MOD match.  depImg_* are CV_16UC1 in millimetres, the rects are
// the template's bounding box in the rendered model image and in the frame.  Throws cv::Exception when a rect leaves
// its image (the reference's cv::Mat ROI does, detection.cpp:43-44).  d_match is accepted and unused, as in the
// reference's live branch (detection.cpp:147, 175-178).
inline void detection(cv::Mat depImg_model_raw, cv::Mat depImg_ref_raw, TCamIntrinsicParam tCamIntrinsic, const cv::Rect_<int> rect_model_raw,
                      cv::Rect_<int> rect_ref_raw, int icp_it_thr, float dist_mean_thr, float dist_diff_thr, cv::Matx33f r_match,
                      cv::Vec3f t_match, float d_match, cv::Vec3f& T_final, cv::Matx33f& R_final) {
  if (depImg_model_raw.type() != CV_16UC1 || depImg_ref_raw.type() != CV_16UC1) throw cv::Exception("detection: depth images must be CV_16UC1");
  if (depImg_model_raw.cols != depImg_ref_raw.cols || depImg_model_raw.rows != depImg_ref_raw.rows) throw cv::Exception("detection: model and reference depth sizes differ");
  const fl_rect_t rm = {rect_model_raw.x, rect_model_raw.y, rect_model_raw.width, rect_model_raw.height};
  const fl_rect_t rr = {rect_ref_raw.x, rect_ref_raw.y, rect_ref_raw.width, rect_ref_raw.height};
  float T[3], R[9];
  const int rc = fl_detection(fealess_b200::icp_handle(), depImg_model_raw.ptr<uint16_t>(0), depImg_model_raw.step, depImg_ref_raw.ptr<uint16_t>(0),
                              depImg_ref_raw.step, depImg_ref_raw.cols, depImg_ref_raw.rows, fealess_b200::intrinsics_of(tCamIntrinsic), rm, rr,
                              icp_it_thr, dist_mean_thr, dist_diff_thr, r_match.val, t_match.val, d_match, T, R);
  fealess_b200::check_status(rc, "detection");
  for (int i = 0; i < 3; ++i) T_final(i) = T[i];
  for (int i = 0; i < 9; ++i) R_final.val[i] = R[i];
}

// icpCloudToCloud_Ex (ICP/ICP.cpp:617-809).  Returns the final mean distance (-1 if a cloud has fewer than 3 points).
inline float icpCloudToCloud_Ex(const std::vector<cv::Vec3f>& pts_ref, const std::vector<cv::Vec3f>& pts_model, cv::Matx33f& R, cv::Vec3f& T,
                                float& px_ratio_match, int icp_it_th = 4, const float dist_mean_thr = 0.0f, const float dist_diff_thr = 0.0f) {
  static_assert(sizeof(cv::Vec3f) == 12, "Vec3f must be three packed floats");
  fl_icp_params_t prm = {icp_it_th, dist_mean_thr, dist_diff_thr};
  fl_icp_result_t r;
  const int rc = fl_icp_cloud_to_cloud_ex(fealess_b200::icp_handle(), pts_ref.empty() ? nullptr : pts_ref[0].val, (int32_t)pts_ref.size(),
                                          pts_model.empty() ? nullptr : pts_model[0].val, (int32_t)pts_model.size(), prm, &r);
  fealess_b200::check_status(rc, "icpCloudToCloud_Ex");
  for (int i = 0; i < 9; ++i) R.val[i] = r.R[i];
  for (int i = 0; i < 3; ++i) T(i) = r.T[i];
  px_ratio_match = r.inlier_ratio;
  return r.dist_mean;
}

namespace cup_d2pc {
// depthTo3d for a CV_16UC1 depth image (ICP/depth_to_3d.cpp:190-221 -> depthTo3dNoMask<float> :99-137 after rescaleDepth
// :244-269): points3d_out = CV_32FC3 in metres, depth 0 -> NaN.  The masked / sparse overloads are not on the path.
inline void depthTo3d(cv::InputArray depth_in, cv::InputArray K_in, cv::OutputArray points3d_out, cv::InputArray mask_in = cv::noArray()) {
  const cv::Mat depth = depth_in.getMat(), K = K_in.getMat();
  if (!mask_in.empty()) throw cv::Exception("depthTo3d: the masked overload is not part of the FEALESS hot path");
  if (depth.type() != CV_16UC1) throw cv::Exception("depthTo3d: only CV_16UC1 depth is on the FEALESS hot path");
  if (K.rows != 3 || K.cols != 3) throw cv::Exception("depthTo3d: K must be 3x3");   // CV_Assert :196
  fl_intrinsics_t k;
  if (K.depth() == CV_32F) { k.fx = K.at<float>(0, 0); k.fy = K.at<float>(1, 1); k.cx = K.at<float>(0, 2); k.cy = K.at<float>(1, 2); }
  else if (K.depth() == CV_64F) { k.fx = (float)K.at<double>(0, 0); k.fy = (float)K.at<double>(1, 1); k.cx = (float)K.at<double>(0, 2); k.cy = (float)K.at<double>(1, 2); }
  else throw cv::Exception("depthTo3d: K must be CV_32F or CV_64F");
#ifdef FEALESS_B200_WITH_OPENCV
  points3d_out.create(depth.rows, depth.cols, CV_32FC3);
  cv::Mat out = points3d_out.getMat();
#else
  if (!points3d_out.mat()) throw cv::Exception("depthTo3d: points3d_out must be a Mat");
  cv::Mat& out = *points3d_out.mat();
  out.create(depth.rows, depth.cols, CV_32FC3);
#endif
  const int rc = fl_depth_to_3d(fealess_b200::icp_handle(), depth.ptr<uint16_t>(0), depth.step, depth.cols, depth.rows, k, out.ptr<float>(0));
  fealess_b200::check_status(rc, "depthTo3d");
}
}  // namespace cup_d2pc

// ICP/obj_data.h:7-17
struct obj_data {
  int match_class;
  float match_sim;
  cv::Mat r;                           // 3x3 rotation
  cv::Mat t;                           // 3x1 translation (CV_32F or CV_64F)
  std::vector<cv::Vec3f> pts_model;
  std::vector<cv::Vec3f> pts_ref;
  float icp_dist;
  bool check_done;
  obj_data() : match_class(0), match_sim(0), icp_dist(0), check_done(false) {}
};

// ICP/pose_result.h:7-106 (the members nonMaximumSuppression and its callers use)
class PoseResult {
 public:
  PoseResult() : R_(9, 0.f), T_(3, 0.f), confidence_(0), object_id_(0) {}
  void set_confidence(float c) { confidence_ = c; }
  void set_object_id(const int& id) { object_id_ = id; }
  void set_R(const cv::Mat& R) { for (int i = 0; i < 9; ++i) R_[i] = elem(R, i / 3, i % 3); }
  void set_T(const cv::Mat& T) { for (int i = 0; i < 3; ++i) T_[i] = T.cols == 1 ? elem(T, i, 0) : elem(T, 0, i); }
  float confidence() const { return confidence_; }
  const int& object_id() const { return object_id_; }
  std::vector<float> R() const { return R_; }
  std::vector<float> T() const { return T_; }
  bool operator==(const PoseResult& p) { return object_id_ == p.object_id_; }
  static float elem(const cv::Mat& m, int r, int c) { return m.depth() == CV_64F ? (float)m.at<double>(r, c) : m.at<float>(r, c); }
 private:
  std::vector<float> R_, T_;
  float confidence_;
  int object_id_;
};

// nonMaximumSuppression (ICP/NMS.cpp:6-39): greedy, order dependent; marks absorbed objects check_done like the reference
// mutates its input.  Objects that arrive with check_done already set are skipped, as in the reference.
inline void nonMaximumSuppression(std::vector<obj_data>& objs, const float th_obj_dist, std::vector<PoseResult>& pose_results) {
  pose_results.clear();
  std::vector<int> live;                                    // indices of the objects the reference's loops would visit
  for (size_t i = 0; i < objs.size(); ++i) if (!objs[i].check_done) live.push_back((int)i);
  if (live.empty()) return;
  const int n = (int)live.size();
  std::vector<float> t3((size_t)n * 3), dist((size_t)n);
  std::vector<int32_t> npts((size_t)n), out_idx((size_t)n);
  for (int k = 0; k < n; ++k) {
    const obj_data& o = objs[live[k]];
    for (int i = 0; i < 3; ++i) t3[3 * k + i] = o.t.cols == 1 ? PoseResult::elem(o.t, i, 0) : PoseResult::elem(o.t, 0, i);
    npts[k] = (int32_t)o.pts_model.size(); dist[k] = o.icp_dist;
  }
  std::vector<uint8_t> absorbed((size_t)n, 0);
  const int cnt = fl_nms_ex(fealess_b200::icp_handle(), t3.data(), npts.data(), dist.data(), n, th_obj_dist, out_idx.data(), absorbed.data());
  if (cnt < 0) fealess_b200::check_status(cnt, "nonMaximumSuppression");
  for (int j = 0; j < n; ++j) if (absorbed[j]) objs[live[j]].check_done = true;   // NMS.cpp:27
  for (int k = 0; k < cnt; ++k) {
    const obj_data& o = objs[live[out_idx[k]]];
    PoseResult pr;
    pr.set_object_id(o.match_class);
    pr.set_confidence(o.match_sim);
    pr.set_R(o.r);
    pr.set_T(o.t);
    pose_results.push_back(pr);
  }
}

#endif  // FEALESS_B200_ICP_HPP
