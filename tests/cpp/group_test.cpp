// Template-sharded Detector::match from C++ with no Python / torch / NCCL in the process: cup_linemod::Detector::useDevices(...) puts
// one handle on every listed GPU (fl_group_*: peer access + the exchange fused into the sort kernel) and must return the list a single
// GPU returns, frame after frame.  usage: group_test <case file> <comma separated device ordinals, e.g. 0,0 or 0,1>
// The case file is the one tests/test_cpp_shim.py writes (frame, templates, threshold, expected matches from the CPU oracle).
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
  if (argc < 3) { std::printf("usage: group_test <case file> <devices>\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::printf("cannot open %s\n", argv[1]); return 2; }
  std::fseek(f, 0, SEEK_END); long sz = std::ftell(f); std::fseek(f, 0, SEEK_SET);
  g_buf.resize((size_t)sz);
  if (std::fread(g_buf.data(), 1, (size_t)sz, f) != (size_t)sz) return 2;
  std::fclose(f);
  std::vector<int> devices;
  for (char* tok = std::strtok(argv[2], ","); tok; tok = std::strtok(nullptr, ",")) devices.push_back(std::atoi(tok));

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

  try {
    cv::Ptr<cup_linemod::Detector> det = cup_linemod::getDefaultLINEMOD();
    det->useDevices(devices, 1024);
    for (int t = 0; t < n_templates; ++t) {
      std::vector<cup_linemod::Template> pyr((size_t)L * M);
      for (int e = 0; e < L * M; ++e) {
        const int32_t* h = &headers[((size_t)t * L * M + e) * 7];
        pyr[e].width = h[0]; pyr[e].height = h[1]; pyr[e].offset_x = h[2]; pyr[e].offset_y = h[3]; pyr[e].pyramid_level = h[4];
        for (int k = 0; k < h[6]; ++k) pyr[e].features.push_back(cup_linemod::Feature(features[3 * (h[5] + k)], features[3 * (h[5] + k) + 1], features[3 * (h[5] + k) + 2]));
      }
      char name[32]; std::snprintf(name, sizeof name, "obj%02d", class_of[t]);
      det->addSyntheticTemplate(pyr, name);
    }
    std::vector<cv::Mat> sources;
    sources.push_back(cv::Mat(H, W, CV_8UC3, bgr.data()));
    sources.push_back(cv::Mat(H, W, CV_16UC1, depth.data()));
    for (int rep = 0; rep < 4; ++rep) {                                   // several frames: both parities of the exchange buffers
      std::vector<cup_linemod::Match> matches;
      std::vector<cv::Mat> quant;
      const int rc = rep == 3 ? det->match(sources, threshold, matches, std::vector<cv::String>(), quant) : det->match(sources, threshold, matches);
      CHECK(rc == 0, "match rc %d", rc);
      CHECK((int)matches.size() == n_expected, "frame %d: %d matches, expected %d", rep, (int)matches.size(), n_expected);
      for (int i = 0; i < n_expected && i < (int)matches.size(); ++i) {
        char name[32]; std::snprintf(name, sizeof name, "obj%02d", expected[i].class_idx);
        CHECK(matches[i].x == expected[i].x && matches[i].y == expected[i].y && matches[i].similarity == expected[i].similarity &&
              matches[i].template_id == expected[i].template_id && matches[i].class_id == name, "frame %d match %d differs", rep, i);
      }
      if (rep == 3) CHECK((int)quant.size() == L * M && quant[0].rows == H && quant[0].cols == W, "quantized images in group mode");
    }
    // one class only
    std::vector<cup_linemod::Match> only;
    std::vector<cv::String> ids; ids.push_back("obj01");
    CHECK(det->match(sources, threshold, only, ids) == 0, "class filter");
    int want = 0; for (int i = 0; i < n_expected; ++i) want += expected[i].class_idx == 1;
    CHECK((int)only.size() == want, "class filter: %d matches, expected %d", (int)only.size(), want);
  } catch (const std::exception& e) {
    std::printf("FAIL exception: %s\n", e.what()); ++g_fail;
  }
  if (g_fail == 0) std::printf("all checks passed (%d devices)\n", (int)devices.size());
  return g_fail ? 1 : 0;
}
