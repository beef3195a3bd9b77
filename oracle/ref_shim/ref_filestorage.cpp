// TEST INFRASTRUCTURE - NOT PRODUCT CODE.  cv::FileStorage stand-in for oracle/_ref (see cvshim.hpp).
// Placeholder bodies: reading / writing template files is not needed by the parity checks yet (templates reach the
// reference's Detector through its own addSyntheticTemplate, linemod.cpp:1624-1630).
#include "cvshim.hpp"

namespace cv {
static void unsupported() { CV_Error(Error::StsNotImplemented, "cvshim: FileStorage is not implemented"); }
FileNode FileNode::operator[](const char*) const { unsupported(); return FileNode(); }
FileNodeIterator FileNode::begin() const { return FileNodeIterator(n_, 0); }
FileNodeIterator FileNode::end() const { return FileNodeIterator(n_, size()); }
FileNode::operator int() const { unsupported(); return 0; }
FileNode::operator float() const { unsupported(); return 0; }
FileNode::operator double() const { unsupported(); return 0; }
FileNode::operator String() const { unsupported(); return String(); }
FileNode FileNodeIterator::operator*() const { unsupported(); return FileNode(); }
FileStorage::FileStorage(const String& filename, int mode) : filename_(filename), mode_(mode), opened_(false), have_key_(false) { unsupported(); }
FileStorage::~FileStorage() {}
void FileStorage::release() {}
void FileStorage::put_string(const std::string&) { unsupported(); }
void FileStorage::put_scalar(const std::string&, bool) { unsupported(); }
FileStorage& operator<<(FileStorage& fs, const char*) { unsupported(); return fs; }
FileStorage& operator<<(FileStorage& fs, const String&) { unsupported(); return fs; }
FileStorage& operator<<(FileStorage& fs, int) { unsupported(); return fs; }
FileStorage& operator<<(FileStorage& fs, float) { unsupported(); return fs; }
FileStorage& operator<<(FileStorage& fs, double) { unsupported(); return fs; }
}  // namespace cv
