// TEST INFRASTRUCTURE — CPU oracle, not the product. See cv_primitives.h for provenance and pins.
#include "cv_primitives.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace cvp {

int round_rne(double v) { return (int)lrint(v); }
int round_rne(float v) { return (int)lrintf(v); }

static inline short sat_s16(int v) { return (short)std::min(32767, std::max(-32768, v)); }

// ---------------------------------------------------------------------------------------------
// resize. Coordinates in double, fractional part in float32, 11-bit weights.
// ---------------------------------------------------------------------------------------------
void resize_linear_coeffs(int dn, int sn, int* ofs, short* c0, short* c1)
{
	const double scale = 1.0 / ((double)dn / sn);
	for (int d = 0; d < dn; d++)
	{
		float f = (float)((d + 0.5) * scale - 0.5);
		int s = (int)std::floor(f);
		f -= (float)s;
		if (s < 0) { s = 0; f = 0.f; }
		if (s + 1 >= sn) { s = sn - 1; f = 0.f; }  // second tap is clamped and carries weight 0
		ofs[d] = s;
		c0[d] = sat_s16(round_rne((1.f - f) * 2048.f));
		c1[d] = sat_s16(round_rne(f * 2048.f));
	}
}

void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                      uint8_t* dst, int dw, int dh, size_t dstep)
{
	std::vector<int> xo(dw), yo(dh);
	std::vector<short> xa0(dw), xa1(dw), ya0(dh), ya1(dh);
	resize_linear_coeffs(dw, sw, xo.data(), xa0.data(), xa1.data());
	resize_linear_coeffs(dh, sh, yo.data(), ya0.data(), ya1.data());

	std::vector<int> row0(dw), row1(dw);
	int cached0 = -1, cached1 = -1;
	auto hpass = [&](int sy, std::vector<int>& row) {
		const uint8_t* s = src + (size_t)sy * sstep;
		for (int dx = 0; dx < dw; dx++)
		{
			const int x0 = xo[dx], x1 = std::min(x0 + 1, sw - 1);
			row[dx] = s[x0] * xa0[dx] + s[x1] * xa1[dx];
		}
	};
	for (int dy = 0; dy < dh; dy++)
	{
		const int sy0 = yo[dy], sy1 = std::min(sy0 + 1, sh - 1);
		if (cached1 == sy0) { std::swap(row0, row1); std::swap(cached0, cached1); }
		if (cached0 != sy0) { hpass(sy0, row0); cached0 = sy0; }
		if (cached1 != sy1) { hpass(sy1, row1); cached1 = sy1; }
		const int b0 = ya0[dy], b1 = ya1[dy];
		uint8_t* d = dst + (size_t)dy * dstep;
		for (int dx = 0; dx < dw; dx++)
		{
			const int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
			d[dx] = (uint8_t)std::min(255, std::max(0, v));
		}
	}
}

// ---------------------------------------------------------------------------------------------
// FAST-9/16. Ring in OpenCV's order; corner iff arc score S > t; response S-1;
// strict 8-neighbour non-max suppression, row-major output.
// ---------------------------------------------------------------------------------------------
static const int RING_DX[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
static const int RING_DY[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

static inline int arc_score_from_diffs(const int* d /*16, centre - ring*/)
{
	// max over the 16 cyclic arcs of 9 of min(d) and of min(-d)
	int best_dark = -256, best_bright = -256;
	for (int k = 0; k < 16; k++)
	{
		int mn = d[k], mx = d[k];
		for (int j = 1; j < 9; j++)
		{
			const int v = d[(k + j) & 15];
			mn = std::min(mn, v);
			mx = std::max(mx, v);
		}
		best_dark = std::max(best_dark, mn);      // ring darker than centre by at least mn on the arc
		best_bright = std::max(best_bright, -mx); // ring brighter than centre by at least -mx on the arc
	}
	return std::max(best_dark, best_bright);
}

int fast9_arc_score(const uint8_t* p, size_t step)
{
	int d[16];
	const int c = p[0];
	for (int k = 0; k < 16; k++)
		d[k] = c - p[(ptrdiff_t)RING_DY[k] * (ptrdiff_t)step + RING_DX[k]];
	return arc_score_from_diffs(d);
}

static inline bool has_arc9(unsigned m)
{
	// m: 16-bit cyclic mask; true iff it holds >= 9 consecutive ones
	auto rot = [](unsigned v, int r) { return ((v << r) | (v >> (16 - r))) & 0xffffu; };
	unsigned a = m & rot(m, 1);
	unsigned b = a & rot(a, 2);
	unsigned c = b & rot(b, 4);
	return (c & rot(m, 8)) != 0;
}

void fast9_16(const uint8_t* img, int w, int h, size_t step, int threshold, bool nms,
              std::vector<FastPoint>& out)
{
	out.clear();
	if (w < 7 || h < 7)
		return;
	threshold = std::min(std::max(threshold, 1), 255);

	ptrdiff_t ofs[16];
	for (int k = 0; k < 16; k++)
		ofs[k] = (ptrdiff_t)RING_DY[k] * (ptrdiff_t)step + RING_DX[k];

	// three rolling score rows (y-1, y, y+1); index 0..w-1, zero where not a corner
	std::vector<int> rows[3];
	for (auto& r : rows) r.assign(w, 0);

	auto score_row = [&](int y, std::vector<int>& row) {
		std::fill(row.begin(), row.end(), 0);
		if (y < 3 || y >= h - 3)
			return;
		const uint8_t* line = img + (size_t)y * step;
		for (int x = 3; x < w - 3; x++)
		{
			const uint8_t* p = line + x;
			const int c = p[0], hi = c + threshold, lo = c - threshold;
			// every arc of 9 holds one pixel of each opposite pair (k, k+8): cheap rejection
			unsigned bright_ok = 1, dark_ok = 1;
			for (int k = 0; k < 8 && (bright_ok | dark_ok); k += 2)
			{
				const int a = p[ofs[k]], b = p[ofs[k + 8]];
				bright_ok &= (unsigned)((a > hi) | (b > hi));
				dark_ok &= (unsigned)((a < lo) | (b < lo));
			}
			if (!(bright_ok | dark_ok))
				continue;
			unsigned mb = 0, md = 0;
			int d[16];
			for (int k = 0; k < 16; k++)
			{
				const int v = p[ofs[k]];
				d[k] = c - v;
				mb |= (unsigned)(v > hi) << k;
				md |= (unsigned)(v < lo) << k;
			}
			if (has_arc9(mb) || has_arc9(md))
				row[x] = arc_score_from_diffs(d) - 1;   // == threshold at minimum, always > 0 for t >= 1
		}
	};

	// Thresholds below 1 are not supported (a corner could then score S-1 == 0); the reference uses 20 and 7.

	auto prev = &rows[0], cur = &rows[1], next = &rows[2];
	score_row(2, *prev);  // all zeros (y < 3)
	score_row(3, *cur);
	for (int y = 3; y < h - 3; y++)
	{
		score_row(y + 1, *next);
		for (int x = 3; x < w - 3; x++)
		{
			const int s = (*cur)[x];
			if (s <= 0)
				continue;
			if (nms)
			{
				if (!(s > (*cur)[x - 1] && s > (*cur)[x + 1] &&
					s > (*prev)[x - 1] && s > (*prev)[x] && s > (*prev)[x + 1] &&
					s > (*next)[x - 1] && s > (*next)[x] && s > (*next)[x + 1]))
					continue;
			}
			out.push_back({ x, y, s });
		}
		auto t = prev; prev = cur; cur = next; next = t;
	}
}

// ---------------------------------------------------------------------------------------------
// GaussianBlur 7x7 sigma 2: 8.8 fixed-point separable kernel, REFLECT_101.
// ---------------------------------------------------------------------------------------------
static const int GK[7] = { 18, 34, 48, 56, 48, 34, 18 };

static inline int reflect101(int i, int n)
{
	if (i < 0) i = -i;
	if (i >= n) i = 2 * n - 2 - i;
	return i;
}

void gauss7x7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
	// horizontal pass for every source row once, then the vertical pass
	std::vector<int> hbuf((size_t)w * h);
	std::vector<int> xi((size_t)w * 7);
	for (int x = 0; x < w; x++)
		for (int j = 0; j < 7; j++)
			xi[(size_t)x * 7 + j] = reflect101(x + j - 3, w);
	for (int y = 0; y < h; y++)
	{
		const uint8_t* s = src + (size_t)y * sstep;
		int* hb = hbuf.data() + (size_t)y * w;
		for (int x = 0; x < w; x++)
		{
			int acc = 0;
			for (int j = 0; j < 7; j++)
				acc += GK[j] * s[xi[(size_t)x * 7 + j]];
			hb[x] = acc;
		}
	}
	for (int y = 0; y < h; y++)
	{
		const int* r[7];
		for (int i = 0; i < 7; i++)
			r[i] = hbuf.data() + (size_t)reflect101(y + i - 3, h) * w;
		uint8_t* d = dst + (size_t)y * dstep;
		for (int x = 0; x < w; x++)
		{
			int acc = 0;
			for (int i = 0; i < 7; i++)
				acc += GK[i] * r[i][x];
			d[x] = (uint8_t)((acc + 32768) >> 16);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// fastAtan2: degree-7 odd polynomial, float32, every operation rounded (no FMA).
// ---------------------------------------------------------------------------------------------
float fast_atan2_deg(float y, float x)
{
	static const float R2D = (float)(180.0 / 3.14159265358979323846);
	static const float p1 = 0.9997878412794807f * R2D;
	static const float p3 = -0.3258083974640975f * R2D;
	static const float p5 = 0.1555786518463281f * R2D;
	static const float p7 = -0.04432655554792128f * R2D;
	const float ax = std::fabs(x), ay = std::fabs(y);
	float a, c, c2;
	if (ax >= ay)
	{
		c = ay / (ax + (float)DBL_EPSILON);
		c2 = c * c;
		a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
	}
	else
	{
		c = ax / (ay + (float)DBL_EPSILON);
		c2 = c * c;
		a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
	}
	if (x < 0) a = 180.f - a;
	if (y < 0) a = 360.f - a;
	return a;
}

// ---------------------------------------------------------------------------------------------
// cvtColor to gray, 8-bit fixed point with 15 fractional bits
// ---------------------------------------------------------------------------------------------
void cvt_gray_u8(const uint8_t* src, int w, int h, size_t sstep, int channels, bool rgb, uint8_t* dst, size_t dstep)
{
	const int ri = rgb ? 0 : 2, bi = rgb ? 2 : 0;
	for (int y = 0; y < h; y++)
	{
		const uint8_t* s = src + (size_t)y * sstep;
		uint8_t* d = dst + (size_t)y * dstep;
		for (int x = 0; x < w; x++, s += channels)
			d[x] = (uint8_t)((s[ri] * 9798 + s[1] * 19235 + s[bi] * 3735 + 16384) >> 15);
	}
}

// ---------------------------------------------------------------------------------------------
// cv::remap, INTER_LINEAR, BORDER_CONSTANT(0), CV_8UC1 source, two CV_32FC1 maps (modules/imgproc/src/imgwarp.cpp): the maps are
// converted to 1/32-pixel fixed point with cvRound (INTER_BITS = 5), the four weights come from a table of products scaled to 2^15
// (exact for bilinear: multiples of 32), taps outside the source read the border value 0, and the sum is rounded by (v + 2^14) >> 15.
// Pinned by tests/golden/primitives.npz (cv2 4.13.0).
// ---------------------------------------------------------------------------------------------
void remap_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, size_t mstep_bytes,
                     uint8_t* dst, int dw, int dh, size_t dstep)
{
	for (int y = 0; y < dh; y++)
	{
		const float* mx = reinterpret_cast<const float*>(reinterpret_cast<const char*>(mapx) + (size_t)y * mstep_bytes);
		const float* my = reinterpret_cast<const float*>(reinterpret_cast<const char*>(mapy) + (size_t)y * mstep_bytes);
		uint8_t* d = dst + (size_t)y * dstep;
		for (int x = 0; x < dw; x++)
		{
			const int sx = round_rne(mx[x] * 32), sy = round_rne(my[x] * 32);
			const int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
			auto tap = [&](int yy, int xx) -> int {
				return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstep + xx] : 0;
			};
			const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
			d[x] = (uint8_t)((tap(iy, ix) * w00 + tap(iy, ix + 1) * w01 + tap(iy + 1, ix) * w10 + tap(iy + 1, ix + 1) * w11 + (1 << 14)) >> 15);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K) (modules/calib3d/src/undistort.dispatch.cpp, cvUndistortPointsInternal):
// everything in double, 5 fixed iterations (TermCriteria(MAX_ITER, 5, 0.01): no epsilon test), coefficients k1 k2 p1 p2 k3 k4 k5 k6
// s1..s4, no tilt; then the new camera matrix P = K is applied and the result is rounded to float. Pinned by tests/golden/primitives.npz.
// ---------------------------------------------------------------------------------------------
void undistort_points(const float* xy, int n, float fx_, float fy_, float cx_, float cy_, const float* dist, int ndist, float* out)
{
	double k[14] = { 0 };
	for (int i = 0; i < ndist && i < 14; i++) k[i] = dist[i];
	const double fx = fx_, fy = fy_, cx = cx_, cy = cy_, ifx = 1. / fx, ify = 1. / fy;
	for (int i = 0; i < n; i++)
	{
		const double u = xy[2 * i], v = xy[2 * i + 1];
		double x = (u - cx) * ifx, y = (v - cy) * ify;
		const double x0 = x, y0 = y;
		for (int j = 0; j < 5; j++)
		{
			const double r2 = x * x + y * y;
			const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
			if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
			const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
			const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
			x = (x0 - deltaX) * icdist;
			y = (y0 - deltaY) * icdist;
		}
		// RR = P * I with P = K: rows (fx 0 cx), (0 fy cy), (0 0 1)
		const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
		out[2 * i] = (float)(xx * ww);
		out[2 * i + 1] = (float)(yy * ww);
	}
}

}  // namespace cvp
