// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// C entry points (flr_*) into the REFERENCE's own LINE-MOD code.  This translation unit pulls in
// /root/reference/linemod/linemod.cpp verbatim (the include below resolves through -I/root/reference/linemod; nothing is
// copied into the repository) so that the file-static stage functions - quantizedOrientations (:230), hysteresisGradient
// (:307), quantizedNormals (:595), spread (:950), computeResponseMaps (:979), linearize (:1060), similarity (:1130),
// similarityLocal (:1226), addSimilarities (:1322) - and the protected Detector::matchClass (:1451) can be called directly,
// next to the public Detector::match (:1356).  Built by oracle/build_ref.py into oracle/_ref/libfl_ref.so; loaded only by
// tests/, __graft_entry__.smoke() and bench.py's reference arm / cpu_baseline leg.
#include "linemod.cpp"

#include <cstdint>

namespace {

struct Probe : public cup_linemod::Detector {
  Probe(const std::vector<cv::Ptr<cup_linemod::Modality> >& mods, const std::vector<int>& T) : cup_linemod::Detector(mods, T) {}
  typedef LinearMemories LMs;
  typedef LinearMemoryPyramid LMPyramid;
  typedef TemplatePyramid TP;
  const TemplatesMap& classes() const { return class_templates; }

  // The front half of Detector::match (linemod.cpp:1385-1416) with the intermediates kept: the calls, their order and their
  // arguments are the reference's; only the storage of the by-products is added.
  void process(const std::vector<cv::Mat>& sources, const std::vector<cv::Mat>& masks, LMPyramid& lm_pyramid, std::vector<cv::Size>& sizes,
               std::vector<cv::Mat>& quantized_out, std::vector<cv::Mat>& spread_out) const {
    std::vector<cv::Ptr<cup_linemod::QuantizedPyramid> > quantizers;
    for (int i = 0; i < (int)modalities.size(); ++i) {
      cv::Mat mask;
      if (!masks.empty()) mask = masks[i];
      quantizers.push_back(modalities[i]->process(sources[i], mask));
    }
    lm_pyramid = LMPyramid(pyramid_levels, std::vector<LMs>(modalities.size(), LMs(8)));
    sizes.clear(); quantized_out.clear(); spread_out.clear();
    for (int l = 0; l < pyramid_levels; ++l) {
      int T = T_at_level[l];
      if (l > 0) for (int i = 0; i < (int)quantizers.size(); ++i) quantizers[i]->pyrDown();
      cv::Mat quantized, spread_quantized;
      std::vector<cv::Mat> response_maps;
      for (int i = 0; i < (int)quantizers.size(); ++i) {
        quantizers[i]->quantize(quantized);
        cup_linemod::spread(quantized, spread_quantized, T);
        cup_linemod::computeResponseMaps(spread_quantized, response_maps);
        for (int j = 0; j < 8; ++j) cup_linemod::linearize(response_maps[j], lm_pyramid[l][i][j], T);
        quantized_out.push_back(quantized.clone());
        spread_out.push_back(spread_quantized.clone());
      }
      sizes.push_back(quantized.size());
    }
  }
  void match_class_raw(const LMPyramid& lm, const std::vector<cv::Size>& sizes, float thr, std::vector<cup_linemod::Match>& out,
                       const cv::String& cls) const {
    TemplatesMap::const_iterator it = class_templates.find(cls);
    if (it != class_templates.end()) matchClass(lm, sizes, thr, out, it->first, it->second);
  }
  int n_modalities() const { return (int)modalities.size(); }
  const std::vector<int>& Ts() const { return T_at_level; }
};

struct Quiet {   // the reference prints progress with std::cout; keep test logs readable
  std::ios_base::iostate saved;
  Quiet() : saved(std::cout.rdstate()) { std::cout.setstate(std::ios_base::failbit); }
  ~Quiet() { std::cout.clear(saved); }
};

}  // namespace

struct flr_match_t { int32_t x, y; float similarity; int32_t class_idx, template_id; };

struct flr_detector {
  cv::Ptr<Probe> det;
  int L, M;
  std::vector<int> kinds;
  std::vector<cv::String> class_names;           // index -> name ("c000000", ... : map order == index order)
  std::map<cv::String, int> class_index;
  std::vector<std::pair<int, int> > tmpl_loc;    // global template index -> (class index, per-class template id)
  // last processed frame
  Probe::LMPyramid lm;
  std::vector<cv::Size> sizes;
  std::vector<cv::Mat> quantized, spread;        // [level * M + modality]
  std::vector<cv::Mat> match_quantized;          // from the real Detector::match
  std::string last_error;
};

static cv::String class_name_of(int idx) { char b[32]; snprintf(b, sizeof b, "c%06d", idx); return cv::String(b); }

extern "C" {

const char* flr_version(void) { return "fl_ref: rlvc/FEALESS linemod.cpp + ICP sources compiled unmodified against oracle/ref_shim"; }

// ---- stage functions: the reference's static functions, called directly ----
int flr_color_quantize(const uint8_t* bgr, int W, int H, float weak_thr, uint8_t* q, float* mag_or_null) {
  try {
    cv::Mat src(H, W, CV_8UC3, (void*)bgr), mag, angle;
    cup_linemod::quantizedOrientations(src, mag, angle, weak_thr);
    for (int r = 0; r < H; ++r) {
      std::memcpy(q + (size_t)r * W, angle.ptr(r), W);
      if (mag_or_null) std::memcpy(mag_or_null + (size_t)r * W, mag.ptr(r), sizeof(float) * W);
    }
    return 0;
  } catch (const cv::Exception&) { return -2; }
}
int flr_depth_quantize(const uint16_t* depth, int W, int H, int dist_thr, int diff_thr, uint8_t* out) {
  try {
    cv::Mat src(H, W, CV_16UC1, (void*)depth), dst;
    cup_linemod::quantizedNormals(src, dst, dist_thr, diff_thr);
    for (int r = 0; r < H; ++r) std::memcpy(out + (size_t)r * W, dst.ptr(r), W);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}
int flr_spread(const uint8_t* q, int W, int H, int T, uint8_t* out) {
  try {
    cv::Mat src(H, W, CV_8UC1, (void*)q), dst;
    cup_linemod::spread(src, dst, T);
    for (int r = 0; r < H; ++r) std::memcpy(out + (size_t)r * W, dst.ptr(r), W);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}
int flr_response_maps(const uint8_t* sp, int W, int H, uint8_t* out8) {
  try {
    cv::Mat src(H, W, CV_8UC1, (void*)sp);
    std::vector<cv::Mat> maps;
    cup_linemod::computeResponseMaps(src, maps);
    for (int i = 0; i < 8; ++i) for (int r = 0; r < H; ++r) std::memcpy(out8 + ((size_t)i * H + r) * W, maps[i].ptr(r), W);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}
int flr_linearize(const uint8_t* resp, int W, int H, int T, uint8_t* out) {
  try {
    cv::Mat src(H, W, CV_8UC1, (void*)resp), lin;
    cup_linemod::linearize(src, lin, T);
    for (int r = 0; r < lin.rows; ++r) std::memcpy(out + (size_t)r * lin.cols, lin.ptr(r), lin.cols);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}

// ---- detector ----
flr_detector* flr_detector_create(int n_levels, const int* T, int n_modalities, const int* modality_kind, float weak_threshold,
                                  int distance_threshold, int difference_threshold) {
  flr_detector* d = new flr_detector;
  d->L = n_levels; d->M = n_modalities;
  std::vector<cv::Ptr<cup_linemod::Modality> > mods;
  for (int i = 0; i < n_modalities; ++i) {
    d->kinds.push_back(modality_kind[i]);
    if (modality_kind[i] == 0) mods.push_back(cv::makePtr<cup_linemod::ColorGradient>(weak_threshold, (size_t)63, 55.0f));
    else mods.push_back(cv::makePtr<cup_linemod::DepthNormal>(distance_threshold, difference_threshold, (size_t)63, 2));
  }
  d->det = cv::makePtr<Probe>(mods, std::vector<int>(T, T + n_levels));
  return d;
}
void flr_detector_destroy(flr_detector* d) { delete d; }
const char* flr_detector_last_error(const flr_detector* d) { return d->last_error.c_str(); }

// headers: 7 ints per (template, level, modality): width,height,offset_x,offset_y,pyramid_level,feature_begin,feature_count;
// features: 3 ints each.  Templates reach the reference through its own addSyntheticTemplate (linemod.cpp:1624-1630).
int flr_detector_set_templates(flr_detector* d, int n_templates, const int32_t* headers, const int32_t* features, int n_features,
                               const int32_t* class_of) {
  // a fresh Detector: the reference has no "clear templates"
  std::vector<cv::Ptr<cup_linemod::Modality> > mods = d->det->getModalities();
  d->det = cv::makePtr<Probe>(mods, d->det->Ts());
  d->class_names.clear(); d->class_index.clear(); d->tmpl_loc.clear();
  const int per = d->L * d->M;
  for (int t = 0; t < n_templates; ++t) {
    int ci = class_of[t];
    if (ci < 0) return -1;
    while ((int)d->class_names.size() <= ci) { d->class_index[class_name_of((int)d->class_names.size())] = (int)d->class_names.size(); d->class_names.push_back(class_name_of((int)d->class_names.size())); }
    std::vector<cup_linemod::Template> tp(per);
    for (int k = 0; k < per; ++k) {
      const int32_t* h = headers + ((size_t)t * per + k) * 7;
      cup_linemod::Template& tt = tp[k];
      tt.width = h[0]; tt.height = h[1]; tt.offset_x = h[2]; tt.offset_y = h[3]; tt.pyramid_level = h[4];
      if (h[5] < 0 || h[6] < 0 || h[5] + h[6] > n_features) return -1;
      for (int f = 0; f < h[6]; ++f) { const int32_t* p = features + (size_t)(h[5] + f) * 3; tt.features.push_back(cup_linemod::Feature(p[0], p[1], p[2])); }
    }
    int tid = d->det->addSyntheticTemplate(tp, d->class_names[ci]);
    d->tmpl_loc.push_back(std::make_pair(ci, tid));
  }
  return 0;
}

static void wrap_sources(const flr_detector* d, const uint8_t* bgr, const uint16_t* depth, int W, int H, const uint8_t* const* masks,
                         std::vector<cv::Mat>& sources, std::vector<cv::Mat>& mask_mats) {
  for (int i = 0; i < d->M; ++i)
    sources.push_back(d->kinds[i] == 0 ? cv::Mat(H, W, CV_8UC3, (void*)bgr) : cv::Mat(H, W, CV_16UC1, (void*)depth));
  if (masks) for (int i = 0; i < d->M; ++i) mask_mats.push_back(masks[i] ? cv::Mat(H, W, CV_8UC1, (void*)masks[i]) : cv::Mat());
}
static void names_of_filter(const flr_detector* d, const int32_t* class_filter, int n_filter, std::vector<cv::String>& ids) {
  for (int i = 0; i < n_filter; ++i)
    ids.push_back(class_filter[i] >= 0 && class_filter[i] < (int)d->class_names.size() ? d->class_names[class_filter[i]] : cv::String("__absent__"));
}
static int export_matches(const flr_detector* d, const std::vector<cup_linemod::Match>& ms, flr_match_t* out, int cap, int* n_total) {
  if (n_total) *n_total = (int)ms.size();
  int n = std::min((int)ms.size(), cap);
  for (int i = 0; i < n; ++i) {
    out[i].x = ms[i].x; out[i].y = ms[i].y; out[i].similarity = ms[i].similarity; out[i].template_id = ms[i].template_id;
    std::map<cv::String, int>::const_iterator it = d->class_index.find(ms[i].class_id);
    out[i].class_idx = it == d->class_index.end() ? -1 : it->second;
  }
  return n;
}

// The REAL cup_linemod::Detector::match (linemod.cpp:1356-1441), final std::sort + std::unique included.
// Returns the number of matches written; -1 = the reference's own -1 (source / mask count); -2 = a CV_Assert fired
// (W % T, H % T, (W*H) % 16, > 63 features ...), message in flr_detector_last_error.
int flr_detector_match(flr_detector* d, const uint8_t* bgr, const uint16_t* depth, int W, int H, const uint8_t* const* masks,
                       float threshold, const int32_t* class_filter, int n_filter, flr_match_t* out, int cap, int* n_total) {
  try {
    Quiet q;
    std::vector<cv::Mat> sources, mask_mats;
    wrap_sources(d, bgr, depth, W, H, masks, sources, mask_mats);
    std::vector<cv::String> ids;
    names_of_filter(d, class_filter, n_filter, ids);
    std::vector<cup_linemod::Match> ms;
    d->match_quantized.clear();
    int rc = d->det->match(sources, threshold, ms, ids, d->match_quantized, mask_mats);
    if (rc != 0) return -1;
    return export_matches(d, ms, out, cap, n_total);
  } catch (const cv::Exception& e) { d->last_error = e.what(); return -2; }
}
// quantised image (level, modality) that the last flr_detector_match returned through `quantized_images`
const uint8_t* flr_detector_match_quantized(const flr_detector* d, int level, int modality, int* W, int* H) {
  size_t k = (size_t)level * d->M + modality;
  if (k >= d->match_quantized.size() || !d->match_quantized[k].isContinuous()) return nullptr;
  *W = d->match_quantized[k].cols; *H = d->match_quantized[k].rows;
  return d->match_quantized[k].data;
}

// front half of match with the by-products kept (Probe::process)
int flr_detector_process(flr_detector* d, const uint8_t* bgr, const uint16_t* depth, int W, int H, const uint8_t* const* masks) {
  try {
    Quiet q;
    std::vector<cv::Mat> sources, mask_mats;
    wrap_sources(d, bgr, depth, W, H, masks, sources, mask_mats);
    d->det->process(sources, mask_mats, d->lm, d->sizes, d->quantized, d->spread);
    return 0;
  } catch (const cv::Exception& e) { d->last_error = e.what(); d->lm.clear(); return -2; }
}
const uint8_t* flr_detector_quantized(const flr_detector* d, int level, int modality, int* W, int* H) {
  const cv::Mat& m = d->quantized[(size_t)level * d->M + modality];
  *W = m.cols; *H = m.rows; return m.data;
}
const uint8_t* flr_detector_spread(const flr_detector* d, int level, int modality) { return d->spread[(size_t)level * d->M + modality].data; }
const uint8_t* flr_detector_lm(const flr_detector* d, int level, int modality, int label, int* rows, int* cols) {
  const cv::Mat& m = d->lm[level][modality][label];
  *rows = m.rows; *cols = m.cols; return m.data;
}
// similarity() of every modality at the coarsest level + addSimilarities (linemod.cpp:1466-1480), u16 [H' * W']
int flr_detector_similarity(const flr_detector* d, int template_idx, uint16_t* out) {
  try {
    if (template_idx < 0 || template_idx >= (int)d->tmpl_loc.size() || d->lm.empty()) return -1;
    const std::vector<cup_linemod::Template>& tp = d->det->getTemplates(d->class_names[d->tmpl_loc[template_idx].first], d->tmpl_loc[template_idx].second);
    std::vector<cv::Mat> sims(d->M);
    int lowest_start = (int)tp.size() - d->M, lowest_T = d->det->Ts().back();
    for (int i = 0; i < d->M; ++i) cup_linemod::similarity(d->lm.back()[i], tp[lowest_start + i], sims[i], d->sizes.back(), lowest_T);
    cv::Mat total;
    cup_linemod::addSimilarities(sims, total);
    for (int r = 0; r < total.rows; ++r) std::memcpy(out + (size_t)r * total.cols, total.ptr(r), sizeof(uint16_t) * total.cols);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}
// similarityLocal of every modality at `level` around (x, y) + addSimilarities (linemod.cpp:1538-1545), u16 [16 * 16]
int flr_detector_similarity_local(const flr_detector* d, int template_idx, int level, int x, int y, uint16_t* out256) {
  try {
    if (template_idx < 0 || template_idx >= (int)d->tmpl_loc.size() || d->lm.empty()) return -1;
    const std::vector<cup_linemod::Template>& tp = d->det->getTemplates(d->class_names[d->tmpl_loc[template_idx].first], d->tmpl_loc[template_idx].second);
    std::vector<cv::Mat> sims(d->M);
    int start = level * d->M, T = d->det->Ts()[level];
    for (int i = 0; i < d->M; ++i) cup_linemod::similarityLocal(d->lm[level][i], tp[start + i], sims[i], d->sizes[level], T, cv::Point(x, y));
    cv::Mat total;
    cup_linemod::addSimilarities(sims, total);
    for (int r = 0; r < 16; ++r) std::memcpy(out256 + r * 16, total.ptr(r), sizeof(uint16_t) * 16);
    return 0;
  } catch (const cv::Exception&) { return -2; }
}
// Detector::matchClass (linemod.cpp:1451-1577) over the processed frame, WITHOUT the final sort / unique: raw emission
// order (class by class in map order, template-major, row-major cells).
int flr_detector_match_raw(flr_detector* d, float threshold, const int32_t* class_filter, int n_filter, flr_match_t* out, int cap, int* n_total) {
  try {
    if (d->lm.empty()) return -1;
    std::vector<cup_linemod::Match> ms;
    if (n_filter == 0) { for (size_t c = 0; c < d->class_names.size(); ++c) d->det->match_class_raw(d->lm, d->sizes, threshold, ms, d->class_names[c]); }
    else { std::vector<cv::String> ids; names_of_filter(d, class_filter, n_filter, ids); for (size_t i = 0; i < ids.size(); ++i) d->det->match_class_raw(d->lm, d->sizes, threshold, ms, ids[i]); }
    return export_matches(d, ms, out, cap, n_total);
  } catch (const cv::Exception& e) { d->last_error = e.what(); return -2; }
}

// Detector::addTemplate (linemod.cpp:1579-1615): the reference's own template extraction on one view.  A scratch detector with the
// same modalities / pyramid receives the template; its pyramid comes back in the layout of flr_detector_set_templates: 7 ints per
// (level, modality) = width, height, offset_x, offset_y, pyramid_level, feature_begin, feature_count; features x, y, label.
// Returns the reference's return value (template id 0, or -1 when a level has too few candidates); bbox = x, y, w, h.
int flr_detector_add_template(flr_detector* d, const uint8_t* bgr, const uint16_t* depth, const uint8_t* mask_or_null, int W, int H,
                              int32_t* headers, int32_t* features, int feature_cap, int* n_features, int32_t bbox[4]) {
  try {
    std::vector<cv::Ptr<cup_linemod::Modality> > mods = d->det->getModalities();
    Probe scratch(mods, d->det->Ts());
    std::vector<cv::Mat> sources;
    for (int i = 0; i < d->M; ++i)
      sources.push_back(d->kinds[i] == 0 ? cv::Mat(H, W, CV_8UC3, const_cast<uint8_t*>(bgr)) : cv::Mat(H, W, CV_16UC1, const_cast<uint16_t*>(depth)));
    cv::Mat mask;
    if (mask_or_null) mask = cv::Mat(H, W, CV_8UC1, const_cast<uint8_t*>(mask_or_null));
    float pose[13] = {0};
    cv::Rect bb;
    const int id = scratch.addTemplate(sources, "train", mask, pose, &bb);
    *n_features = 0;
    if (id < 0) return id;
    const std::vector<cup_linemod::Template>& tp = scratch.getTemplates("train", id);
    int nf = 0;
    for (size_t k = 0; k < tp.size(); ++k) {
      int32_t* h = headers + k * 7;
      h[0] = tp[k].width; h[1] = tp[k].height; h[2] = tp[k].offset_x; h[3] = tp[k].offset_y; h[4] = tp[k].pyramid_level; h[5] = nf; h[6] = (int)tp[k].features.size();
      for (size_t j = 0; j < tp[k].features.size(); ++j, ++nf) {
        if (nf >= feature_cap) return -3;
        features[3 * nf] = tp[k].features[j].x; features[3 * nf + 1] = tp[k].features[j].y; features[3 * nf + 2] = tp[k].features[j].label;
      }
    }
    *n_features = nf;
    bbox[0] = bb.x; bbox[1] = bb.y; bbox[2] = bb.width; bbox[3] = bb.height;
    return id;
  } catch (const cv::Exception& e) { d->last_error = e.what(); return -2; }
}
// the two OpenCV primitives only template training calls (linemod.cpp:466, 753, 765), for pinning the stand-ins on cv2
void flr_prim_erode3(const uint8_t* src, int W, int H, int iterations, uint8_t* out) {
  cv::Mat s(H, W, CV_8UC1, const_cast<uint8_t*>(src)), o;
  cv::erode(s, o, cv::Mat(), cv::Point(-1, -1), iterations, cv::BORDER_REPLICATE);
  std::memcpy(out, o.ptr(0), (size_t)W * H);
}
void flr_prim_distance_c3(const uint8_t* src, int W, int H, float* out) {
  cv::Mat s(H, W, CV_8UC1, const_cast<uint8_t*>(src)), o;
  cv::distanceTransform(s, o, cv::DIST_C, 3);
  std::memcpy(out, o.ptr(0), (size_t)W * H * sizeof(float));
}

}  // extern "C"
