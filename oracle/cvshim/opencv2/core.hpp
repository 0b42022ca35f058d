// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// Minimal stand-in for the slice of OpenCV the reference hot path touches, so that
// /root/reference/src/ORBextractor.cc (and the matcher line ranges) compile UNMODIFIED without an
// OpenCV install. It also lets the drop-in host class
// (orb_slam2_refactored_b200/csrc/host) build here; with a real OpenCV on the include path this
// directory is simply not used.
//
// Matrices are byte buffers with an element size: 8-bit with 1/3/4 channels and 32-bit float exist. The four arithmetic primitives forward to the pinned
// restatements in oracle/cv_primitives.cc.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_8UC4 24
#define CV_32F 5
#define CV_32FC1 5
#define CV_PI 3.1415926535897932384626433832795

namespace cv {

namespace Error { enum Code { StsError = -2, StsAssert = -215 }; }

// the constructors real OpenCV has (core.hpp: Exception() and Exception(code, err, func, file, line)) and no others
class Exception : public std::runtime_error
{
public:
	Exception() : std::runtime_error("") {}
	Exception(int code_, const std::string& err_, const std::string& func_, const std::string& file_, int line_)
		: std::runtime_error(file_ + ":" + std::to_string(line_) + ": error: (" + std::to_string(code_) + ") " + err_ + " in function '" + func_ + "'"),
		  code(code_), err(err_), func(func_), file(file_), line(line_) {}
	int code = 0;
	std::string err, func, file;
	int line = 0;
};

}  // namespace cv

#define CV_Assert(expr) do { if (!(expr)) throw cv::Exception(cv::Error::StsAssert, #expr, __func__, __FILE__, __LINE__); } while (0)

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }

namespace cv {

template <class T> struct Point_
{
	T x, y;
	Point_() : x(0), y(0) {}
	Point_(T x_, T y_) : x(x_), y(y_) {}
	Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
// OpenCV's types.hpp: a - b = Point_<T>(saturate_cast<T>(a.x - b.x), saturate_cast<T>(a.y - b.y))
template <class T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>((T)(a.x - b.x), (T)(a.y - b.y)); }
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect
{
	int x, y, width, height;
	Rect() : x(0), y(0), width(0), height(0) {}
	Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

// cv::Matx as far as the camera-pose headers use it (include/CameraPose.h, include/CameraProjection.h). The arithmetic follows
// modules/core/include/opencv2/core/matx.hpp of the pinned OpenCV: a product accumulates s = 0; s += a(i,k)*b(k,j) for k ascending
// in the element type, a sum is element-wise, unary minus scales by -1.
template <class T, int M, int N> struct Matx
{
	T val[M * N];
	Matx() { for (int i = 0; i < M * N; i++) val[i] = T(0); }
	Matx(T v0, T v1, T v2) { static_assert(M * N == 3, "3 values"); val[0] = v0; val[1] = v1; val[2] = v2; }
	Matx(T v0, T v1, T v2, T v3, T v4, T v5, T v6, T v7, T v8)
	{
		static_assert(M * N == 9, "9 values");
		val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; val[4] = v4; val[5] = v5; val[6] = v6; val[7] = v7; val[8] = v8;
	}
	static Matx zeros() { return Matx(); }
	static Matx eye() { Matx m; for (int i = 0; i < (M < N ? M : N); i++) m.val[i * N + i] = T(1); return m; }
	T& operator()(int i, int j) { return val[i * N + j]; }
	const T& operator()(int i, int j) const { return val[i * N + j]; }
	T& operator()(int i) { static_assert(M == 1 || N == 1, "vector"); return val[i]; }
	const T& operator()(int i) const { static_assert(M == 1 || N == 1, "vector"); return val[i]; }
	Matx<T, N, M> t() const { Matx<T, N, M> r; for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) r.val[j * M + i] = val[i * N + j]; return r; }
	T dot(const Matx& o) const { T s = 0; for (int i = 0; i < M * N; i++) s += val[i] * o.val[i]; return s; }
};
template <class T, int M, int L, int N> inline Matx<T, M, N> operator*(const Matx<T, M, L>& a, const Matx<T, L, N>& b)
{
	Matx<T, M, N> r;
	for (int i = 0; i < M; i++)
		for (int j = 0; j < N; j++)
		{
			T s = 0;
			for (int k = 0; k < L; k++) s += a(i, k) * b(k, j);
			r.val[i * N + j] = s;
		}
	return r;
}
template <class T, int M, int N> inline Matx<T, M, N> operator+(const Matx<T, M, N>& a, const Matx<T, M, N>& b)
{
	Matx<T, M, N> r;
	for (int i = 0; i < M * N; i++) r.val[i] = a.val[i] + b.val[i];
	return r;
}
template <class T, int M, int N> inline Matx<T, M, N> operator-(const Matx<T, M, N>& a, const Matx<T, M, N>& b)
{
	Matx<T, M, N> r;
	for (int i = 0; i < M * N; i++) r.val[i] = a.val[i] - b.val[i];
	return r;
}
template <class T, int M, int N> inline Matx<T, M, N> operator-(const Matx<T, M, N>& a)
{
	Matx<T, M, N> r;
	for (int i = 0; i < M * N; i++) r.val[i] = a.val[i] * -1;
	return r;
}
// scalar products: Matx(a, alpha, Matx_ScaleOp): val[i] = saturate_cast<T>(a.val[i] * alpha)
template <class T, int M, int N> inline Matx<T, M, N> operator*(const Matx<T, M, N>& a, float alpha)
{
	Matx<T, M, N> r;
	for (int i = 0; i < M * N; i++) r.val[i] = (T)(a.val[i] * alpha);
	return r;
}
template <class T, int M, int N> inline Matx<T, M, N> operator*(float alpha, const Matx<T, M, N>& a) { return a * alpha; }
typedef Matx<float, 3, 1> Matx31f;
typedef Matx<float, 3, 3> Matx33f;
// cv::norm(Matx) = std::sqrt(normL2Sqr<_Tp, double>(val, m*n)): squares accumulated in double (matx.hpp, base.hpp)
template <class T, int M, int N> inline double norm(const Matx<T, M, N>& a)
{
	double s = 0;
	for (int i = 0; i < M * N; i++) s += (double)a.val[i] * (double)a.val[i];
	return std::sqrt(s);
}

struct KeyPoint
{
	Point2f pt;
	float size;
	float angle;
	float response;
	int octave;
	int class_id;
	KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
	KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
		: pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "KeyPoint must mirror cv::KeyPoint (7 x 4 bytes)");

enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };

class Mat
{
public:
	int rows, cols;
	size_t step;
	uchar* data;
	int type_;

	Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
	Mat(int r, int c, int type) : Mat() { create(r, c, type); }
	// external (non-owning) buffer, like cv::Mat(rows, cols, type, void*, step)
	Mat(int r, int c, int type, void* ext, size_t step_ = 0)
		: rows(r), cols(c), step(step_ ? step_ : (size_t)c * esz(type)), data((uchar*)ext), type_(type) {}

	static size_t esz(int type) { return ((type & 7) == 5 ? 4 : 1) * (size_t)((type >> 3) + 1); }
	int type() const { return type_; }
	int channels() const { return (type_ >> 3) + 1; }
	size_t elemSize() const { return esz(type_); }
	bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
	size_t step1() const { return step / ((type_ & 7) == 5 ? 4 : 1); }
	bool isContinuous() const { return step == (size_t)cols * elemSize(); }

	void create(int r, int c, int type)
	{
		if (r == rows && c == cols && type == type_ && data && owner_ && step == (size_t)c * esz(type))
			return;
		owner_ = std::make_shared<std::vector<uchar>>((size_t)r * c * esz(type));
		rows = r; cols = c; type_ = type; step = (size_t)c * esz(type); data = owner_->data();
	}
	void create(Size s, int type) { create(s.height, s.width, type); }
	void release() { owner_.reset(); rows = cols = 0; step = 0; data = nullptr; }
	void setTo(int v)
	{
		for (int y = 0; y < rows; y++)
			std::memset(data + (size_t)y * step, v, (size_t)cols * elemSize());
	}
	void copyTo(Mat& dst) const
	{
		dst.create(rows, cols, type_);
		for (int y = 0; y < rows; y++)
			std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
	}
	Mat clone() const { Mat m; copyTo(m); return m; }

	Mat operator()(const Range& rr, const Range& cr) const
	{
		Mat m;
		m.owner_ = owner_;
		m.type_ = type_;
		m.rows = rr.end - rr.start; m.cols = cr.end - cr.start; m.step = step;
		m.data = data + (size_t)rr.start * step + (size_t)cr.start * elemSize();
		return m;
	}
	Mat operator()(const Rect& r) const { return (*this)(Range(r.y, r.y + r.height), Range(r.x, r.x + r.width)); }
	Mat row(int r) const { return (*this)(Range(r, r + 1), Range(0, cols)); }

	uchar* ptr(int r = 0) { return data + (size_t)r * step; }
	const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
	template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
	template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
	template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
	template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }

private:
	std::shared_ptr<std::vector<uchar>> owner_;
};

template <class T> class Mat_ : public Mat
{
public:
	Mat_() {}
	Mat_(const Mat& m) : Mat(m) {}
	Mat_(int r, int c) : Mat(r, c, sizeof(T) == 4 ? CV_32F : CV_8U) {}
	static Mat_ eye(int r, int c)
	{
		Mat_ m(r, c);
		m.setTo(0);
		for (int i = 0; i < (r < c ? r : c); i++) m(i, i) = T(1);
		return m;
	}
	T& operator()(int i) { return reinterpret_cast<T*>(this->data)[i]; }                 // vector-like access, continuous matrices
	const T& operator()(int i) const { return reinterpret_cast<const T*>(this->data)[i]; }
	T& operator()(int y, int x) { return this->template at<T>(y, x); }
	const T& operator()(int y, int x) const { return this->template at<T>(y, x); }
};
typedef Mat_<uchar> Mat1b;

// pinned restatements, defined in oracle/cvshim_impl.cc on top of oracle/cv_primitives.cc
void resize(const Mat& src, Mat& dst, Size dsize);
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
float fastAtan2(float y, float x);
enum { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7, COLOR_BGRA2GRAY = 10, COLOR_RGBA2GRAY = 11 };
void cvtColor(const Mat& src, Mat& dst, int code);
typedef Mat_<float> Mat1f;
// cv::undistortPoints as src/System.cc:165 calls it: points in place, K 3x3 CV_32F, distortion coefficients CV_32F, no R, P = K
void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& K, const Mat& distCoeffs, const Mat& R, const Mat& P);
enum { INTER_LINEAR = 1 };
// cv::remap as Examples/Stereo/stereo_euroc.cc:100-101 calls it: 8-bit single channel, two CV_32F maps, INTER_LINEAR, default border
void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation);

}  // namespace cv
