/* Stand-in for the reference's cvheaders.h (Lib/TLibCommon/cvheaders.h:1-5), which pulls in
 * OpenCV 2.x.  OpenCV is not in this image and none of the cv:: code executes with the fork's
 * default `Naive` decision model (tools_YS.cpp:5,37), so this header only has to make the
 * reference's sources COMPILE under g++.  It is test infrastructure for oracle/_ref and is
 * never part of the product.  Nothing here computes a number on the hot path. */
#ifndef CUCD_REF_CVHEADERS_STUB_H
#define CUCD_REF_CVHEADERS_STUB_H
#include <string>
#include <vector>
#include <cstdlib>

#define CV_8U 0
#define CV_32FC1 5
#define CV_ROW_SAMPLE 1
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS 2

namespace cv {
typedef std::string String;
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
class Mat {
public:
  int rows, cols, dims;
  Mat() : rows(0), cols(0), dims(0) {}
  Mat(int r, int c, int) : rows(r), cols(c), dims(2) {}
  Mat(int, const int*, int) : rows(0), cols(0), dims(0) {}
  template <class T> explicit Mat(const std::vector<T>&) : rows(0), cols(0), dims(0) {}
  static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
  template <class T> T& at(int = 0, int = 0, int = 0) { static T dummy; std::abort(); return dummy; }
  template <class T> T& at(const int*) { static T dummy; std::abort(); return dummy; }
  Mat colRange(int, int) const { return Mat(); }
  Mat rowRange(int, int) const { return Mat(); }
  Mat reshape(int, int = 0) const { return Mat(); }
  Mat reshape(int, int, const int*) const { return Mat(); }
  Mat clone() const { return *this; }
  Mat t() const { return *this; }
  bool empty() const { return true; }
  int channels() const { return 1; }
  int total() const { return 0; }
  Size size() const { return Size(cols, rows); }
  void copyTo(Mat&) const {}
  void convertTo(Mat&, int) const {}
  template <class T> void push_back(const T&) {}
  Mat& setTo(double) { return *this; }
};
typedef Mat MatND;
struct SVM { enum { C_SVC = 100, NU_SVC, ONE_CLASS, EPS_SVR, NU_SVR }; enum { LINEAR = 0, POLY, RBF, SIGMOID }; };
struct TermCriteria { TermCriteria(int = 0, int = 0, double = 0) {} };
}  // namespace cv
using namespace cv;

struct CvTermCriteria { int type, max_iter; double epsilon; };
inline CvTermCriteria cvTermCriteria(int t, int m, double e) { CvTermCriteria c = {t, m, e}; return c; }
struct CvSVMParams {
  int svm_type, kernel_type; double degree, gamma, coef0, C, nu, p; CvTermCriteria term_crit;
  CvSVMParams() : svm_type(0), kernel_type(0), degree(0), gamma(0), coef0(0), C(1), nu(0), p(0) {}
};
struct CvRTParams {
  CvRTParams() {}
  CvRTParams(int, int, float, bool, int, const float*, bool, int, int, float, int) {}
};
class CvSVM {
public:
  bool train(const Mat&, const Mat&, const Mat& = Mat(), const Mat& = Mat(), CvSVMParams = CvSVMParams()) { std::abort(); return false; }
  float predict(const Mat&, bool = false) const { std::abort(); return 0.f; }
  void save(const char*, const char* = 0) const {}
  void load(const char*, const char* = 0) {}
};
class CvRTrees {
public:
  bool train(const Mat&, int, const Mat&, const Mat& = Mat(), const Mat& = Mat(), const Mat& = Mat(), const Mat& = Mat(), CvRTParams = CvRTParams()) { std::abort(); return false; }
  float predict(const Mat&, const Mat& = Mat()) const { std::abort(); return 0.f; }
  void save(const char*, const char* = 0) const {}
  void load(const char*, const char* = 0) {}
};
#endif
