// cup_linemod::Detector::addTemplate of the C++ mirror (include/fealess_b200/linemod.hpp), called the way a template generator calls
// the reference's (sources, class id, object mask, pose, &bounding_box), against the pyramid the Python side expects (which the GPU
// tests tie to the reference's own addTemplate).  Then the trained template is matched on its own view.  Exit code 0 = all passed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fealess_b200/linemod.hpp"

static std::vector<unsigned char> g_buf;
static size_t g_pos = 0;
template <typename T> static T rd() { T v; std::memcpy(&v, &g_buf[g_pos], sizeof(T)); g_pos += sizeof(T); return v; }
template <typename T> static std::vector<T> rdv(size_t n) { std::vector<T> v(n); if (n) std::memcpy(v.data(), &g_buf[g_pos], n * sizeof(T)); g_pos += n * sizeof(T); return v; }
static int g_fail = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: train_test <case file>\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::printf("cannot open %s\n", argv[1]); return 2; }
  std::fseek(f, 0, SEEK_END); long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  g_buf.resize((size_t)sz);
  if (std::fread(g_buf.data(), 1, (size_t)sz, f) != (size_t)sz) return 2;
  std::fclose(f);
  const int W = rd<int32_t>(), H = rd<int32_t>();
  std::vector<uint8_t> bgr = rdv<uint8_t>((size_t)W * H * 3);
  std::vector<uint16_t> depth = rdv<uint16_t>((size_t)W * H);
  std::vector<uint8_t> mask = rdv<uint8_t>((size_t)W * H), tiny = rdv<uint8_t>((size_t)W * H);
  const int n_entries = rd<int32_t>();
  std::vector<int32_t> headers = rdv<int32_t>((size_t)n_entries * 7);
  const int n_features = rd<int32_t>();
  std::vector<int32_t> features = rdv<int32_t>((size_t)n_features * 3);
  std::vector<int32_t> bbox = rdv<int32_t>(4);

  cv::Ptr<cup_linemod::Detector> det = cup_linemod::getDefaultLINEMOD();
  std::vector<cv::Mat> sources;
  sources.push_back(cv::Mat(H, W, CV_8UC3, bgr.data()));
  sources.push_back(cv::Mat(H, W, CV_16UC1, depth.data()));
  float pose[13]; for (int i = 0; i < 13; ++i) pose[i] = (float)i;
  cv::Rect bb;
  CHECK(det->addTemplate(sources, "obj", cv::Mat(H, W, CV_8UC1, tiny.data()), pose, &bb) == -1, "a mask too small for 63 features must return -1");
  CHECK(det->numTemplates() == 0 && det->numPoseInfos() == 0, "a failed addTemplate must add nothing");
  const int id = det->addTemplate(sources, "obj", cv::Mat(H, W, CV_8UC1, mask.data()), pose, &bb);
  CHECK(id == 0 && det->numTemplates("obj") == 1 && det->numPoseInfos() == 1, "template id %d", id);
  CHECK(bb.x == bbox[0] && bb.y == bbox[1] && bb.width == bbox[2] && bb.height == bbox[3], "bounding box %d %d %d %d", bb.x, bb.y, bb.width, bb.height);
  const std::vector<cup_linemod::Template>& pyr = det->getTemplates("obj", 0);
  CHECK((int)pyr.size() == n_entries, "pyramid entries %d", (int)pyr.size());
  for (int e = 0; e < n_entries && e < (int)pyr.size(); ++e) {
    const int32_t* h = &headers[(size_t)e * 7];
    CHECK(pyr[e].width == h[0] && pyr[e].height == h[1] && pyr[e].offset_x == h[2] && pyr[e].offset_y == h[3] && pyr[e].pyramid_level == h[4] && (int)pyr[e].features.size() == h[6],
          "header of entry %d", e);
    for (int k = 0; k < h[6] && k < (int)pyr[e].features.size(); ++k) {
      const cup_linemod::Feature& ft = pyr[e].features[k];
      const int32_t* g = &features[(size_t)(h[5] + k) * 3];
      if (ft.x != g[0] || ft.y != g[1] || ft.label != g[2]) { CHECK(false, "feature %d of entry %d: %d %d %d != %d %d %d", k, e, ft.x, ft.y, ft.label, g[0], g[1], g[2]); break; }
    }
  }
  std::vector<cup_linemod::Match> matches;
  CHECK(det->match(sources, 90.f, matches) == 0 && !matches.empty(), "the trained template must match its own view");
  if (!matches.empty()) CHECK(std::abs(matches[0].x - bb.x) <= 5 && std::abs(matches[0].y - bb.y) <= 5 && matches[0].class_id == "obj", "best match at %d %d", matches[0].x, matches[0].y);
  if (g_fail == 0) std::printf("all checks passed\n");
  return g_fail ? 1 : 0;
}
