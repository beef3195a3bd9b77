// fealess_b200/linemod_io.hpp - readLinemod / writeLinemod for the C++ host-side mirror: the reference's single-file template
// database (reference: linemod/linemod_if.cpp:36-66; Detector::read / write linemod/linemod.cpp:1681-1708; readClass /
// writeClass :1710-1786; Template::read / write :98-129; Feature::write :38-41; ColorGradient / DepthNormal read / write
// :536-558, :857-876).  This is what CObjRecoLmICP::AddObj calls first (CadReco/obj_reco_lmicp.cpp:67-74), so with it the
// mirror covers AddObj -> Recognition end to end.
//
// The reference parses with cv::FileStorage.  This header carries its own reader and writer for the YAML dialect FileStorage
// emits (block maps and sequences by indentation, flow sequences / maps that may wrap over lines, quoted strings, the
// "%YAML:1.0" directive), so it needs no OpenCV: it builds against cv_min.hpp here and against the real SDK in a FEALESS
// tree, and the files are interchangeable with cv::FileStorage in both directions (tests/test_cpp_io.py checks both against
// OpenCV's own parser).  XML files (which FileStorage also accepts) are handled by the Python binding only.
//
// Behaviour kept from the reference: a file that cannot be opened yields an EMPTY detector (numClasses() == 0 - AddObj turns
// that into ERROR_OPEN_FILE_FAILED); the CV_Asserts of readClass (modality names, pyramid_levels, duplicate class, template ids
// out of order) throw cv::Exception; template_pose rows go to ONE flat list shared by all classes (:1745-1746, SURVEY.md
// A.6 iv) and writeClass writes TemplatePoseInfo[i] for the i-th template of EVERY class (:1776).
#ifndef FEALESS_B200_LINEMOD_IO_HPP
#define FEALESS_B200_LINEMOD_IO_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "linemod.hpp"

namespace fealess_b200 {
namespace fsio {

// one node of the parsed file (the role of cv::FileNode)
struct Node {
  enum Kind { NONE, SCALAR, SEQ, MAP };
  Kind kind;
  std::string text;               // SCALAR
  std::vector<std::string> keys;  // MAP: keys[i] names items[i]
  std::vector<Node> items;        // SEQ elements / MAP values
  Node() : kind(NONE) {}
  bool empty() const { return kind == NONE; }
  size_t size() const { return kind == SCALAR ? 1 : items.size(); }
  const Node& operator[](const std::string& key) const {
    static const Node none;
    if (kind == MAP) for (size_t i = 0; i < keys.size(); ++i) if (keys[i] == key) return items[i];
    return none;
  }
  // cv::FileNode's conversions: a missing node reads as 0 / ""
  double real() const {
    if (kind != SCALAR) return 0.0;
    const std::string& s = text;
    if (s == ".Inf" || s == "+.Inf" || s == ".inf") return HUGE_VAL;
    if (s == "-.Inf" || s == "-.inf") return -HUGE_VAL;
    if (s == ".NaN" || s == ".nan") return std::nan("");
    return std::strtod(s.c_str(), nullptr);
  }
  int integer() const { const double v = real(); return (int)(v < 0 ? std::ceil(v - 0.5) : std::floor(v + 0.5)); }   // cvRound-like for "5."
  const std::string& string() const { static const std::string none; return kind == SCALAR ? text : none; }
};

namespace detail {
struct Line { int indent; std::string text; };

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
  return s.substr(a, b - a);
}
// net bracket depth of s outside quoted strings
inline int bracket_balance(const std::string& s) {
  int depth = 0; char quote = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    const char c = s[i];
    if (quote) { if (c == '\\' && quote == '"') ++i; else if (c == quote) quote = 0; continue; }
    if (c == '"' || c == '\'') quote = c;
    else if (c == '[' || c == '{') ++depth;
    else if (c == ']' || c == '}') --depth;
  }
  return depth;
}
inline std::string unquote(const std::string& raw) {
  const std::string s = trim(raw);
  if (s.size() >= 2 && s[0] == '"' && s[s.size() - 1] == '"') {
    std::string out;
    for (size_t i = 1; i + 1 < s.size(); ++i) {
      if (s[i] == '\\' && i + 2 < s.size()) {
        const char e = s[++i];
        out += e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e;
      } else out += s[i];
    }
    return out;
  }
  if (s.size() >= 2 && s[0] == '\'' && s[s.size() - 1] == '\'') return s.substr(1, s.size() - 2);
  return s;
}
// position of the ':' that separates a block-map key from its value (followed by a blank or the end of the line), or npos
inline size_t key_colon(const std::string& s) {
  char quote = 0;
  for (size_t i = 0; i < s.size(); ++i) {
    const char c = s[i];
    if (quote) { if (c == '\\' && quote == '"') ++i; else if (c == quote) quote = 0; continue; }
    if (c == '"' || c == '\'') { quote = c; continue; }
    if (c == '[' || c == '{') return std::string::npos;
    if (c == ':' && (i + 1 == s.size() || s[i + 1] == ' ')) return i;
  }
  return std::string::npos;
}

class Parser {
 public:
  explicit Parser(std::istream& in) {
    std::string raw;
    while (std::getline(in, raw)) {
      size_t ind = 0;
      while (ind < raw.size() && raw[ind] == ' ') ++ind;
      std::string t = trim(raw);
      if (t.empty() || t[0] == '#' || t[0] == '%' || t == "---" || t == "...") continue;
      // a flow collection may wrap over several lines: glue them until the brackets balance
      int depth = bracket_balance(t);
      while (depth > 0 && std::getline(in, raw)) { const std::string more = trim(raw); t += ' '; t += more; depth += bracket_balance(more); }
      Line l; l.indent = (int)ind; l.text = t;
      lines_.push_back(l);
    }
  }
  Node parse() {
    pos_ = 0;
    if (lines_.empty()) return Node();
    return block(lines_[0].indent);
  }

 private:
  static bool is_item(const std::string& t) { return t == "-" || (t.size() > 1 && t[0] == '-' && t[1] == ' '); }
  static std::string strip_tag(const std::string& v) {           // "!!opencv-matrix" style tags are skipped
    if (v.size() > 1 && v[0] == '!') { const size_t sp = v.find(' '); return sp == std::string::npos ? std::string() : trim(v.substr(sp)); }
    return v;
  }
  Node value(const std::string& v) {
    Node n;
    if (!v.empty() && (v[0] == '[' || v[0] == '{')) { size_t p = 0; return flow(v, p); }
    n.kind = Node::SCALAR; n.text = unquote(v);
    return n;
  }
  Node block(int indent) {
    Node n;
    if (pos_ >= lines_.size()) return n;
    if (is_item(lines_[pos_].text)) {
      n.kind = Node::SEQ;
      while (pos_ < lines_.size() && lines_[pos_].indent == indent && is_item(lines_[pos_].text)) {
        const std::string rest = trim(lines_[pos_].text.substr(1));
        if (rest.empty()) {                                       // "-" alone: the element is the deeper block that follows
          ++pos_;
          if (pos_ < lines_.size() && lines_[pos_].indent > indent) n.items.push_back(block(lines_[pos_].indent));
          else n.items.push_back(Node());
        } else if (rest[0] != '[' && rest[0] != '{' && key_colon(rest) != std::string::npos) {
          // "- key: value": the element is a map whose first entry sits on the dash's line
          const int inner = indent + (int)(lines_[pos_].text.size() - rest.size());
          lines_[pos_].indent = inner; lines_[pos_].text = rest;
          n.items.push_back(block(inner));
        } else {
          n.items.push_back(value(strip_tag(rest)));
          ++pos_;
        }
      }
      return n;
    }
    n.kind = Node::MAP;
    while (pos_ < lines_.size() && lines_[pos_].indent == indent && !is_item(lines_[pos_].text)) {
      const std::string& t = lines_[pos_].text;
      const size_t c = key_colon(t);
      if (c == std::string::npos) throw cv::Exception("readLinemod: cannot parse line '" + t + "'");
      const std::string key = unquote(t.substr(0, c));
      const std::string v = strip_tag(trim(t.substr(c + 1)));
      ++pos_;
      n.keys.push_back(key);
      if (!v.empty()) n.items.push_back(value(v));
      else if (pos_ < lines_.size() && (lines_[pos_].indent > indent || (lines_[pos_].indent == indent && is_item(lines_[pos_].text))))
        n.items.push_back(block(lines_[pos_].indent));
      else n.items.push_back(Node());
    }
    return n;
  }
  static void skip_blanks(const std::string& s, size_t& p) { while (p < s.size() && (s[p] == ' ' || s[p] == '\t')) ++p; }
  // one flow scalar starting at p: quoted, or up to the next , ] } (a ':' ends it too inside a flow map)
  static std::string flow_scalar(const std::string& s, size_t& p, bool in_map_key) {
    skip_blanks(s, p);
    const size_t b = p;
    if (p < s.size() && (s[p] == '"' || s[p] == '\'')) {
      const char q = s[p++];
      while (p < s.size() && s[p] != q) { if (s[p] == '\\' && q == '"') ++p; ++p; }
      if (p < s.size()) ++p;
      return unquote(s.substr(b, p - b));
    }
    while (p < s.size() && s[p] != ',' && s[p] != ']' && s[p] != '}' && !(in_map_key && s[p] == ':')) ++p;
    return trim(s.substr(b, p - b));
  }
  Node flow(const std::string& s, size_t& p) {
    Node n;
    skip_blanks(s, p);
    if (p >= s.size()) return n;
    if (s[p] == '[' || s[p] == '{') {
      const bool is_map = s[p] == '{';
      const char close = is_map ? '}' : ']';
      n.kind = is_map ? Node::MAP : Node::SEQ;
      ++p;
      if (p < s.size() && s[p] == ':') ++p;                       // OpenCV's "[:" compact-sequence spelling
      for (;;) {
        skip_blanks(s, p);
        if (p >= s.size()) throw cv::Exception("readLinemod: unterminated flow collection");
        if (s[p] == close) { ++p; break; }
        if (s[p] == ',') { ++p; continue; }
        if (is_map) {
          n.keys.push_back(flow_scalar(s, p, true));
          skip_blanks(s, p);
          if (p < s.size() && s[p] == ':') ++p;
        }
        n.items.push_back(flow(s, p));
      }
      return n;
    }
    n.kind = Node::SCALAR;
    n.text = flow_scalar(s, p, false);
    return n;
  }

  std::vector<Line> lines_;
  size_t pos_ = 0;
};

// ---- writer: the layout cv::FileStorage produces for the same calls (3-space indentation, "-" on its own line before a
// map element, "[ a, b ]" flow sequences wrapped at ~80 columns) ----
class Writer {
 public:
  explicit Writer(std::ostream& out) : out_(out), depth_(0) { out_ << "%YAML:1.0\n---\n"; }
  void key_int(const char* k, int v) { pad(); out_ << k << ": " << v << "\n"; }
  void key_real(const char* k, double v) { pad(); out_ << k << ": " << real(v) << "\n"; }
  void key_str(const char* k, const std::string& v) { pad(); out_ << k << ": " << str(v) << "\n"; }
  void begin_seq(const char* k) { pad(); out_ << k << ":\n"; ++depth_; }
  void end_seq() { --depth_; }
  void begin_map_item() { pad(); out_ << "-\n"; ++depth_; }
  void end_map_item() { --depth_; }
  // "key: [ a, b, ... ]" or, with k == nullptr, the sequence element "- [ a, b, ... ]"
  void flow_seq(const char* k, const std::vector<std::string>& v) {
    std::string line(3 * (size_t)depth_, ' ');
    line += k ? std::string(k) + ": [" : std::string("- [");
    const std::string cont(3 * (size_t)depth_ + 4, ' ');
    for (size_t i = 0; i < v.size(); ++i) {
      const std::string tok = " " + v[i] + (i + 1 < v.size() ? "," : "");
      if (line.size() + tok.size() > 78 && line.size() > cont.size()) { out_ << line << "\n"; line = cont; line += tok.substr(1); }
      else line += tok;
    }
    out_ << line << " ]\n";
  }
  static std::string real(double v) {
    if (std::isnan(v)) return ".NaN";
    if (std::isinf(v)) return v < 0 ? "-.Inf" : ".Inf";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", v);                   // round-trips the float -> double value exactly
    std::string s(buf);
    if (s.find_first_of(".eEn") == std::string::npos) s += '.';   // "10." keeps the scalar a real for FileStorage
    return s;
  }
  static std::string str(const std::string& v) {                  // FileStorage quotes anything that is not a plain word
    bool plain = !v.empty() && !(v[0] >= '0' && v[0] <= '9') && v[0] != '-' && v[0] != '+' && v[0] != '.';
    for (size_t i = 0; i < v.size() && plain; ++i) {
      const char c = v[i];
      plain = (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' || c == '.';
    }
    if (plain) return v;
    std::string q = "\"";
    for (size_t i = 0; i < v.size(); ++i) { if (v[i] == '"' || v[i] == '\\') q += '\\'; q += v[i]; }
    return q + "\"";
  }

 private:
  void pad() { for (int i = 0; i < 3 * depth_; ++i) out_ << ' '; }
  std::ostream& out_;
  int depth_;
};

inline std::string itos(int v) { return std::to_string(v); }
}  // namespace detail

inline Node parse_stream(std::istream& in) { detail::Parser p(in); return p.parse(); }
inline Node parse_file(const std::string& filename) {
  std::ifstream in(filename.c_str());
  if (!in) return Node();                                         // FileStorage that failed to open: root() is an empty node
  return parse_stream(in);
}

// Modality::create(const FileNode&) (linemod.cpp:208-223) + ColorGradient::read / DepthNormal::read
inline cv::Ptr<cup_linemod::Modality> read_modality(const Node& fn) {
  const std::string type = fn["type"].string();
  if (type == "ColorGradient") {
    return cv::Ptr<cup_linemod::Modality>(new cup_linemod::ColorGradient((float)fn["weak_threshold"].real(), (size_t)fn["num_features"].integer(),
                                                                         (float)fn["strong_threshold"].real()));
  }
  if (type == "DepthNormal") {
    return cv::Ptr<cup_linemod::Modality>(new cup_linemod::DepthNormal(fn["distance_threshold"].integer(), fn["difference_threshold"].integer(),
                                                                       (size_t)fn["num_features"].integer(), fn["extract_threshold"].integer()));
  }
  throw cv::Exception("readLinemod: unknown modality type '" + type + "'");   // the reference would dereference an empty Ptr
}

// Template::read (linemod.cpp:98-113)
inline cup_linemod::Template read_template(const Node& fn) {
  cup_linemod::Template t;
  t.width = fn["width"].integer(); t.height = fn["height"].integer();
  t.offset_x = fn["offset_x"].integer(); t.offset_y = fn["offset_y"].integer();
  t.pyramid_level = fn["pyramid_level"].integer();
  const Node& feats = fn["features"];
  t.features.resize(feats.size());
  for (size_t i = 0; i < feats.size(); ++i) {
    const Node& f = feats.items[i];                               // Feature::read (:30-36): [x, y, label]
    if (f.kind != Node::SEQ || f.items.size() != 3) throw cv::Exception("readLinemod: a feature is not an [x, y, label] triple");
    t.features[i] = cup_linemod::Feature(f.items[0].integer(), f.items[1].integer(), f.items[2].integer());
  }
  return t;
}

// Detector::readClass (linemod.cpp:1710-1762)
inline std::string read_class(cup_linemod::Detector& det, const Node& fn) {
  const std::vector<cv::Ptr<cup_linemod::Modality> >& mods = det.getModalities();
  const Node& mod_fn = fn["modalities"];
  if (mod_fn.size() != mods.size()) throw cv::Exception("readClass: mod_fn.size() == modalities.size()");
  for (size_t i = 0; i < mods.size(); ++i)
    if (mods[i]->name() != mod_fn.items[i].string()) throw cv::Exception("readClass: modalities[i]->name() == (String)(*mod_it)");
  if (fn["pyramid_levels"].integer() != det.pyramidLevels()) throw cv::Exception("readClass: (int)fn[\"pyramid_levels\"] == pyramid_levels");
  const std::string class_id = fn["class_id"].string();
  if (det.numTemplates(class_id) != 0) throw cv::Exception("readClass: class '" + class_id + "' is already registered");
  const Node& tps = fn["template_pyramids"];
  det.notePoseOffset(class_id);                                   // (mirror-only bookkeeping for getPoseInfo(class_id, template_id))
  for (size_t e = 0; e < tps.size(); ++e) {
    const Node& tp = tps.items[e];
    if (tp["template_id"].integer() != (int)e) throw cv::Exception("readClass: template_id == expected_id");
    const Node& pose_fn = tp["template_pose"];
    float pose[13] = {0};
    for (size_t k = 0; k < pose_fn.size() && k < 13; ++k) pose[k] = (float)pose_fn.items[k].real();
    det.addPoseInfo(pose);                                        // flat list, like TemplatePoseInfo.push_back(ff_fn) :1749
    const Node& templates_fn = tp["templates"];
    std::vector<cup_linemod::Template> pyramid(templates_fn.size());
    for (size_t j = 0; j < templates_fn.size(); ++j) pyramid[j] = read_template(templates_fn.items[j]);
    det.addSyntheticTemplate(pyramid, class_id);
  }
  return class_id;
}

}  // namespace fsio
}  // namespace fealess_b200

// readLinemod (linemod_if.cpp:36-47): Detector::read on the root, then readClass for every entry of "classes"
inline cv::Ptr<cup_linemod::Detector> readLinemod(const std::string& filename) {
  namespace io = fealess_b200::fsio;
  const io::Node root = io::parse_file(filename);
  if (root.empty()) return cv::Ptr<cup_linemod::Detector>(new cup_linemod::Detector());   // unreadable file: empty detector, numClasses() == 0
  std::vector<int> T;
  const io::Node& t_fn = root["T"];
  for (size_t i = 0; i < t_fn.size(); ++i) T.push_back(t_fn.items[i].integer());
  if (root["pyramid_levels"].integer() != (int)T.size()) throw cv::Exception("readLinemod: pyramid_levels does not match the length of T");
  std::vector<cv::Ptr<cup_linemod::Modality> > mods;
  const io::Node& mods_fn = root["modalities"];
  for (size_t i = 0; i < mods_fn.size(); ++i) mods.push_back(io::read_modality(mods_fn.items[i]));
  cv::Ptr<cup_linemod::Detector> det(new cup_linemod::Detector(mods, T));
  const io::Node& classes = root["classes"];
  for (size_t i = 0; i < classes.size(); ++i) io::read_class(*det, classes.items[i]);
  return det;
}

// writeLinemod (linemod_if.cpp:49-66): Detector::write, then one map per class (writeClass)
inline void writeLinemod(const cv::Ptr<cup_linemod::Detector>& detector, const std::string& filename) {
  namespace io = fealess_b200::fsio;
  using io::detail::itos;
  std::ofstream out(filename.c_str());
  if (!out) throw cv::Exception("writeLinemod: cannot open '" + filename + "' for writing");
  io::detail::Writer w(out);
  const int L = detector->pyramidLevels();
  w.key_int("pyramid_levels", L);
  std::vector<std::string> toks;
  for (int l = 0; l < L; ++l) toks.push_back(itos(detector->getT(l)));
  w.flow_seq("T", toks);
  const std::vector<cv::Ptr<cup_linemod::Modality> >& mods = detector->getModalities();
  w.begin_seq("modalities");
  for (size_t m = 0; m < mods.size(); ++m) {
    w.begin_map_item();
    w.key_str("type", mods[m]->name());
    if (const cup_linemod::ColorGradient* cg = dynamic_cast<const cup_linemod::ColorGradient*>(mods[m].get())) {
      w.key_real("weak_threshold", cg->weak_threshold); w.key_int("num_features", (int)cg->num_features); w.key_real("strong_threshold", cg->strong_threshold);
    } else if (const cup_linemod::DepthNormal* dn = dynamic_cast<const cup_linemod::DepthNormal*>(mods[m].get())) {
      w.key_int("distance_threshold", dn->distance_threshold); w.key_int("difference_threshold", dn->difference_threshold);
      w.key_int("num_features", (int)dn->num_features); w.key_int("extract_threshold", dn->extract_threshold);
    }
    w.end_map_item();
  }
  w.end_seq();
  const std::vector<cv::String> ids = detector->classIds();
  w.begin_seq("classes");
  for (size_t c = 0; c < ids.size(); ++c) {
    w.begin_map_item();
    w.key_str("class_id", ids[c]);
    toks.clear();
    for (size_t m = 0; m < mods.size(); ++m) toks.push_back(mods[m]->name());
    w.flow_seq("modalities", toks);
    w.key_int("pyramid_levels", L);
    w.begin_seq("template_pyramids");
    const int n = detector->numTemplates(ids[c]);
    for (int i = 0; i < n; ++i) {
      w.begin_map_item();
      w.key_int("template_id", i);
      toks.clear();
      if (i < detector->numPoseInfos()) { const std::vector<float> pose = detector->getPoseInfo(i); for (size_t k = 0; k < pose.size(); ++k) toks.push_back(io::detail::Writer::real(pose[k])); }
      w.flow_seq("template_pose", toks);                           // TemplatePoseInfo[i], :1776
      w.begin_seq("templates");
      const std::vector<cup_linemod::Template>& tp = detector->getTemplates(ids[c], i);
      for (size_t j = 0; j < tp.size(); ++j) {
        w.begin_map_item();
        w.key_int("width", tp[j].width); w.key_int("height", tp[j].height);
        w.key_int("offset_x", tp[j].offset_x); w.key_int("offset_y", tp[j].offset_y);
        w.key_int("pyramid_level", tp[j].pyramid_level);
        w.begin_seq("features");
        for (size_t k = 0; k < tp[j].features.size(); ++k) {
          toks.clear();
          toks.push_back(itos(tp[j].features[k].x)); toks.push_back(itos(tp[j].features[k].y)); toks.push_back(itos(tp[j].features[k].label));
          w.flow_seq(nullptr, toks);
        }
        w.end_seq();
        w.end_map_item();
      }
      w.end_seq();
      w.end_map_item();
    }
    w.end_seq();
    w.end_map_item();
  }
  w.end_seq();
}

#endif  // FEALESS_B200_LINEMOD_IO_HPP
