// TEST INFRASTRUCTURE — CPU oracle, not the product.
//
// Restatements of the four un-vendored OpenCV primitives the reference hot path calls
// (cv::resize / cv::FAST / cv::GaussianBlur / cv::fastAtan2) plus cvRound. OpenCV is a
// find_package dependency of the reference (CMakeLists.txt:20, version unpinned); the pin chosen for
// this build is OpenCV 4.13.0 as observed through python cv2. Every function here
// is checked bit-for-bit against cv2 4.13.0 by tests/golden/make_golden.py (fixtures committed).
//
// Call sites in the reference these stand in for:
//   cv::resize       src/ORBextractor.cc:468
//   cv::FAST         src/ORBextractor.cc:527,530
//   cv::GaussianBlur src/ORBextractor.cc:799
//   cv::fastAtan2    src/ORBextractor.cc:100
//   cv::cvtColor     src/System.cc:136 (RGB/BGR/RGBA/BGRA -> GRAY)
//   cvRound          src/ORBextractor.cc:78,109,113-114,466-467,482,547,709
//
// Build flags are normative: -O2 -ffp-contract=off, no -march=native, no -ffast-math.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cvp {

struct FastPoint { int x, y, score; };

// round-half-to-even, as cvRound (lrint in the default rounding mode)
int round_rne(double v);
int round_rne(float v);

// cv::resize(src, dst, Size(dw,dh)), 8UC1, INTER_LINEAR (fixed-point, 11-bit coefficients)
void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                      uint8_t* dst, int dw, int dh, size_t dstep);

// Coefficient tables of the resize above, exposed so the CUDA host side can be checked against them.
// ofs[d] = first source tap, c0/c1 = int16 weights (sum 2048).
void resize_linear_coeffs(int dn, int sn, int* ofs, short* c0, short* c1);

// cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16. Output order is row-major.
void fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, bool nms,
              std::vector<FastPoint>& out);

// Threshold-independent arc score S of one pixel (corner iff S > t, response = S-1). Needs a 3 px margin.
int fast9_arc_score(const uint8_t* p, size_t step);

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101), 8UC1, whole (non-sub) matrix
void gauss7x7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);

// cv::cvtColor(src, dst, COLOR_{RGB,BGR,RGBA,BGRA}2GRAY), 8-bit: (R*9798 + G*19235 + B*3735 + 16384) >> 15
// (OpenCV 4.13.0, verified on 5 M random pixels, IPP on and off). channels = 3 or 4; rgb = first channel is R.
void cvt_gray_u8(const uint8_t* src, int w, int h, size_t sstep, int channels, bool rgb, uint8_t* dst, size_t dstep);
void undistort_points(const float* xy, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist, float* out);
void remap_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, size_t mstep_bytes,
                     uint8_t* dst, int dw, int dh, size_t dstep);

// cv::fastAtan2(y, x) in degrees [0,360)
float fast_atan2_deg(float y, float x);

}  // namespace cvp
