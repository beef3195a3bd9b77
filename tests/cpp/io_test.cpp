// Driver of tests/test_cpp_io.py: readLinemod / writeLinemod of the C++ mirror (include/fealess_b200/linemod_io.hpp), no GPU.
//   io_test dump <file>          canonical text dump of what readLinemod built
//   io_test rewrite <in> <out>   readLinemod(in) -> writeLinemod(out)
#include <cstdio>
#include <cstring>
#include <string>

#include "fealess_b200_compat/linemod_if.h"

static int dump(const std::string& path) {
  cv::Ptr<cup_linemod::Detector> det = readLinemod(path);
  std::printf("levels %d\nT", det->pyramidLevels());
  for (int l = 0; l < det->pyramidLevels(); ++l) std::printf(" %d", det->getT(l));
  std::printf("\n");
  const std::vector<cv::Ptr<cup_linemod::Modality> >& mods = det->getModalities();
  for (size_t m = 0; m < mods.size(); ++m) {
    if (const cup_linemod::ColorGradient* cg = dynamic_cast<const cup_linemod::ColorGradient*>(mods[m].get()))
      std::printf("modality ColorGradient %.9g %d %.9g\n", cg->weak_threshold, (int)cg->num_features, cg->strong_threshold);
    else if (const cup_linemod::DepthNormal* dn = dynamic_cast<const cup_linemod::DepthNormal*>(mods[m].get()))
      std::printf("modality DepthNormal %d %d %d %d\n", dn->distance_threshold, dn->difference_threshold, (int)dn->num_features, dn->extract_threshold);
  }
  std::printf("classes %d templates %d poses %d\n", det->numClasses(), det->numTemplates(), det->numPoseInfos());
  const std::vector<cv::String> ids = det->classIds();
  for (size_t c = 0; c < ids.size(); ++c) {
    std::printf("class %s %d\n", ids[c].c_str(), det->numTemplates(ids[c]));
    for (int t = 0; t < det->numTemplates(ids[c]); ++t) {
      const std::vector<cup_linemod::Template>& tp = det->getTemplates(ids[c], t);
      for (size_t j = 0; j < tp.size(); ++j) {
        std::printf("t %d %zu %d %d %d %d %d %zu", t, j, tp[j].width, tp[j].height, tp[j].offset_x, tp[j].offset_y, tp[j].pyramid_level, tp[j].features.size());
        for (size_t k = 0; k < tp[j].features.size(); ++k) std::printf(" %d %d %d", tp[j].features[k].x, tp[j].features[k].y, tp[j].features[k].label);
        std::printf("\n");
      }
    }
  }
  for (int i = 0; i < det->numPoseInfos(); ++i) {
    const std::vector<float> p = det->getPoseInfo(i);
    std::printf("pose %d", i);
    for (size_t k = 0; k < p.size(); ++k) std::printf(" %.9g", p[k]);
    std::printf("\n");
  }
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc == 3 && !std::strcmp(argv[1], "dump")) return dump(argv[2]);
    if (argc == 4 && !std::strcmp(argv[1], "rewrite")) { writeLinemod(readLinemod(argv[2]), argv[3]); return 0; }
  } catch (const cv::Exception& e) {
    std::printf("cv::Exception: %s\n", e.what());
    return 3;
  }
  std::printf("usage: io_test dump <file> | rewrite <in> <out>\n");
  return 2;
}
