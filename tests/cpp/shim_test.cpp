// Drives the C++ mirror of the reference interface (include/fealess_b200/linemod.hpp, icp.hpp) the way
// CadReco/obj_reco_lmicp.cpp drives the reference (Recognition, :86-204): build a Detector, add templates, match a frame,
// refine the best match with detection(), run NMS - and compares every result with the expectations the Python side
// computed with the CPU oracle (tests/test_cpp_shim.py writes the case file).  Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "fealess_b200/icp.hpp"
#include "fealess_b200/linemod.hpp"

static std::vector<unsigned char> g_buf;
static size_t g_pos = 0;
template <typename T> static T rd() { T v; std::memcpy(&v, &g_buf[g_pos], sizeof(T)); g_pos += sizeof(T); return v; }
template <typename T> static std::vector<T> rdv(size_t n) { std::vector<T> v(n); if (n) std::memcpy(v.data(), &g_buf[g_pos], n * sizeof(T)); g_pos += n * sizeof(T); return v; }

static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: shim_test <case file>\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::printf("cannot open %s\n", argv[1]); return 2; }
  std::fseek(f, 0, SEEK_END); long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  g_buf.resize((size_t)sz);
  if (std::fread(g_buf.data(), 1, (size_t)sz, f) != (size_t)sz) return 2;
  std::fclose(f);

  // ---- LINE-MOD ------------------------------------------------------------------------------
  const int W = rd<int32_t>(), H = rd<int32_t>();
  std::vector<uint8_t> bgr = rdv<uint8_t>((size_t)W * H * 3);
  std::vector<uint16_t> depth = rdv<uint16_t>((size_t)W * H);
  const int L = rd<int32_t>(), M = rd<int32_t>();
  std::vector<int32_t> T = rdv<int32_t>(L);
  const int n_templates = rd<int32_t>();
  std::vector<int32_t> headers = rdv<int32_t>((size_t)n_templates * L * M * 7);
  const int n_features = rd<int32_t>();
  std::vector<int32_t> features = rdv<int32_t>((size_t)n_features * 3);
  std::vector<int32_t> class_of = rdv<int32_t>(n_templates);
  const float threshold = rd<float>();
  const int n_expected = rd<int32_t>();
  std::vector<fl_match_t> expected = rdv<fl_match_t>(n_expected);

  cv::Ptr<cup_linemod::Detector> det = cup_linemod::getDefaultLINEMOD();
  CHECK(det->pyramidLevels() == L && (int)det->getModalities().size() == M && det->getT(0) == T[0] && det->getT(1) == T[1], "default detector shape");
  for (int t = 0; t < n_templates; ++t) {
    std::vector<cup_linemod::Template> pyr((size_t)L * M);
    for (int e = 0; e < L * M; ++e) {
      const int32_t* h = &headers[((size_t)t * L * M + e) * 7];
      pyr[e].width = h[0]; pyr[e].height = h[1]; pyr[e].offset_x = h[2]; pyr[e].offset_y = h[3]; pyr[e].pyramid_level = h[4];
      for (int k = 0; k < h[6]; ++k) pyr[e].features.push_back(cup_linemod::Feature(features[3 * (h[5] + k)], features[3 * (h[5] + k) + 1], features[3 * (h[5] + k) + 2]));
    }
    char name[32]; std::snprintf(name, sizeof name, "obj%02d", class_of[t]);
    det->addSyntheticTemplate(pyr, name);
    float pose[13]; for (int i = 0; i < 13; ++i) pose[i] = (float)(t * 13 + i);
    det->addPoseInfo(pose);
  }
  CHECK(det->numTemplates() == n_templates, "numTemplates %d != %d", det->numTemplates(), n_templates);
  CHECK(det->getPoseInfo(3)[12] == 3 * 13 + 12, "pose info");

  cv::Mat rgb(H, W, CV_8UC3, bgr.data()), dep(H, W, CV_16UC1, depth.data());
  std::vector<cv::Mat> sources; sources.push_back(rgb); sources.push_back(dep);
  std::vector<cup_linemod::Match> matches;
  std::vector<cv::String> class_ids;
  std::vector<cv::Mat> quantized_images;
  int rc = det->match(sources, threshold, matches, class_ids, quantized_images);   // the call Recognition makes (:101)
  CHECK(rc == 0, "match rc %d", rc);
  CHECK((int)matches.size() == n_expected, "match count %d != %d", (int)matches.size(), n_expected);
  CHECK((int)quantized_images.size() == L * M && quantized_images[0].rows == H && quantized_images[2].rows == H / 2, "quantized images");
  for (int i = 0; i < n_expected && i < (int)matches.size(); ++i) {
    char name[32]; std::snprintf(name, sizeof name, "obj%02d", expected[i].class_idx);
    const cup_linemod::Match& m = matches[i];
    CHECK(m.x == expected[i].x && m.y == expected[i].y && m.similarity == expected[i].similarity && m.class_id == name && m.template_id == expected[i].template_id,
          "match %d: (%d,%d,%.6f,%s,%d) != (%d,%d,%.6f,%s,%d)", i, m.x, m.y, m.similarity, m.class_id.c_str(), m.template_id, expected[i].x, expected[i].y,
          expected[i].similarity, name, expected[i].template_id);
  }
  // argument checks that return -1 in the reference (linemod.cpp:1364-1378)
  std::vector<cv::Mat> one; one.push_back(rgb);
  CHECK(det->match(one, threshold, matches) == -1, "sources/modalities mismatch must return -1");
  std::vector<cv::Mat> bad_masks(1);
  CHECK(det->match(sources, threshold, matches, class_ids, cv::noArray(), bad_masks) == -1, "mask count mismatch must return -1");
  // class filter: unknown ids are ignored
  std::vector<cv::String> only; only.push_back("no_such_class");
  CHECK(det->match(sources, threshold, matches, only) == 0 && matches.empty(), "unknown class id must match nothing");
  // CV_Assert analogue: 642 columns are not divisible by T
  bool threw = false;
  try {
    std::vector<uint8_t> b2((size_t)642 * H * 3); std::vector<uint16_t> d2((size_t)642 * H);
    std::vector<cv::Mat> s2; s2.push_back(cv::Mat(H, 642, CV_8UC3, b2.data())); s2.push_back(cv::Mat(H, 642, CV_16UC1, d2.data()));
    det->match(s2, threshold, matches);
  } catch (const cv::Exception&) { threw = true; }
  CHECK(threw, "W %% T != 0 must throw cv::Exception");

  // ---- detection() ------------------------------------------------------------------------------
  std::vector<uint16_t> model_depth = rdv<uint16_t>((size_t)W * H), ref_depth = rdv<uint16_t>((size_t)W * H);
  std::vector<int32_t> rm = rdv<int32_t>(4), rr = rdv<int32_t>(4);
  std::vector<float> r0 = rdv<float>(9), t0 = rdv<float>(3), Texp = rdv<float>(3), Rexp = rdv<float>(9);
  TCamIntrinsicParam K; K.nWidth = W; K.nHeight = H; K.dFx = 608; K.dFy = 608; K.dCx = 320; K.dCy = 240;
  cv::Matx33f r_match, R_final; cv::Vec3f t_match, T_final;
  for (int i = 0; i < 9; ++i) r_match.val[i] = r0[i];
  for (int i = 0; i < 3; ++i) t_match(i) = t0[i];
  detection(cv::Mat(H, W, CV_16UC1, model_depth.data()), cv::Mat(H, W, CV_16UC1, ref_depth.data()), K, cv::Rect_<int>(rm[0], rm[1], rm[2], rm[3]),
            cv::Rect_<int>(rr[0], rr[1], rr[2], rr[3]), 10, 0.5f, 0.01f, r_match, t_match, 0.f, T_final, R_final);
  for (int i = 0; i < 3; ++i) CHECK(std::fabs(T_final(i) - Texp[i]) < 0.1f, "T_final[%d] %.5f vs %.5f (tolerance 1e-4 m)", i, T_final(i), Texp[i]);
  {   // |skew(Rexp^T R)| / 2 < 1e-4 rad
    double D[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { D[3 * i + j] = 0; for (int k = 0; k < 3; ++k) D[3 * i + j] += (double)Rexp[3 * k + i] * R_final.val[3 * k + j]; }
    const double e = std::sqrt((D[7] - D[5]) * (D[7] - D[5]) + (D[2] - D[6]) * (D[2] - D[6]) + (D[3] - D[1]) * (D[3] - D[1])) / 2;
    CHECK(e < 1e-4, "rotation error %.3g rad", e);
  }
  threw = false;
  try { detection(cv::Mat(H, W, CV_16UC1, model_depth.data()), cv::Mat(H, W, CV_16UC1, ref_depth.data()), K, cv::Rect_<int>(rm[0], rm[1], rm[2], rm[3]),
                  cv::Rect_<int>(W - 10, rr[1], rr[2], rr[3]), 10, 0.5f, 0.01f, r_match, t_match, 0.f, T_final, R_final); }
  catch (const cv::Exception&) { threw = true; }
  CHECK(threw, "rect outside the frame must throw");

  // ---- depthTo3d ----------------------------------------------------------------------------------
  cv::Mat Kmat(3, 3, CV_32FC1); std::memset(Kmat.data, 0, 36);
  Kmat.at<float>(0, 0) = 608; Kmat.at<float>(1, 1) = 608; Kmat.at<float>(0, 2) = 320; Kmat.at<float>(1, 2) = 240; Kmat.at<float>(2, 2) = 1;
  cv::Mat pts;
  cup_d2pc::depthTo3d(cv::Mat(H, W, CV_16UC1, ref_depth.data()), Kmat, pts);
  CHECK(pts.rows == H && pts.cols == W && pts.type() == CV_32FC3, "depthTo3d output shape");
  int bad = 0;
  for (int v = 0; v < H; v += 7) for (int u = 0; u < W; u += 5) {
    const uint16_t d = ref_depth[(size_t)v * W + u];
    const float* p = pts.ptr<float>(v) + 3 * u;
    if (d == 0) { if (!(p[2] != p[2])) ++bad; continue; }
    const float z = (float)d * (float)(1 / 1000.0), x = (float)(u - 320.0f) * (1.0f / 608.0f) * z;
    if (p[2] != z || std::fabs(p[0] - x) > 1e-6f) ++bad;
  }
  CHECK(bad == 0, "depthTo3d: %d sampled pixels differ", bad);

  // ---- nonMaximumSuppression ------------------------------------------------------------------------
  const int n_obj = rd<int32_t>();
  std::vector<float> t3 = rdv<float>((size_t)n_obj * 3);
  std::vector<int32_t> npts = rdv<int32_t>(n_obj);
  std::vector<float> dist = rdv<float>(n_obj);
  const int n_out = rd<int32_t>();
  std::vector<int32_t> out_idx = rdv<int32_t>(n_out);
  std::vector<obj_data> objs((size_t)n_obj);
  for (int i = 0; i < n_obj; ++i) {
    objs[i].match_class = i; objs[i].match_sim = 90.f - i; objs[i].icp_dist = dist[i];
    objs[i].r = cv::Mat::zeros(3, 3, CV_32FC1); objs[i].r.at<float>(0, 0) = objs[i].r.at<float>(1, 1) = objs[i].r.at<float>(2, 2) = 1.f;
    objs[i].t = cv::Mat(3, 1, CV_32FC1);
    for (int k = 0; k < 3; ++k) objs[i].t.at<float>(k, 0) = t3[3 * i + k];
    objs[i].pts_model.resize((size_t)npts[i]);
  }
  std::vector<PoseResult> poses;
  nonMaximumSuppression(objs, 30.0f, poses);
  CHECK((int)poses.size() == n_out, "NMS count %d != %d", (int)poses.size(), n_out);
  for (int i = 0; i < n_out && i < (int)poses.size(); ++i) CHECK(poses[i].object_id() == out_idx[i], "NMS[%d] object %d != %d", i, poses[i].object_id(), out_idx[i]);
  int n_done = 0; for (int i = 0; i < n_obj; ++i) n_done += objs[i].check_done ? 1 : 0;
  CHECK(n_done == n_obj - n_out, "check_done flags: %d absorbed, expected %d", n_done, n_obj - n_out);

  std::printf(g_fail ? "shim_test: %d check(s) FAILED\n" : "shim_test: all checks passed (%d matches, launches %lld)\n", g_fail ? g_fail : (int)n_expected,
              (long long)fl_launch_count(det->handle()));
  return g_fail ? 1 : 0;
}
