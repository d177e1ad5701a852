// TEST INFRASTRUCTURE shim (oracle/_ref build only): the handful of OpenCV core
// types the reference node touches. OpenCV C++ headers are absent in this image.
// cv::Mat keeps OpenCV's semantics that matter here: copies are shallow
// (ref-counted), clone() is deep, at<T>(row,col) is row-major.
// imread() does not decode PNGs: the host (Python cv2) decodes + resizes the map
// exactly as grid_map.cpp:30-36 would and hands the bytes over through
// cvshim::set_next_image(); resize() then sees equal sizes and copies.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <stdexcept>
namespace cv {
typedef unsigned char uchar;
template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point_<int> Point2i; typedef Point_<double> Point2d; typedef Point2i Point;
typedef Point3_<double> Point3d;
struct Vec3b { uchar v[3]; Vec3b() { v[0] = v[1] = v[2] = 0; } Vec3b(uchar a, uchar b, uchar c) { v[0] = a; v[1] = b; v[2] = c; } };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
#define CV_8UC(n) (n)
#define CV_8UC1 1
#define CV_8UC3 3
enum { IMREAD_GRAYSCALE = 0, COLOR_GRAY2BGR = 8 };
class Mat {
public:
    int rows = 0, cols = 0, ch = 1;
    std::shared_ptr<std::vector<uchar>> buf;
    uchar* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(const Mat& m, const Rect&) { *this = m; }  // ROI only used by the GUI (never enabled)
    void create(int r, int c, int type) {
        rows = r; cols = c; ch = type;
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c * ch);
        data = buf->data();
    }
    Mat clone() const {
        Mat m; m.rows = rows; m.cols = cols; m.ch = ch;
        if (buf) { m.buf = std::make_shared<std::vector<uchar>>(*buf); m.data = m.buf->data(); }
        return m;
    }
    void copyTo(Mat& o) const { o = clone(); }
    Size size() const { return Size(cols, rows); }
    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + ((size_t)r * cols + c) * sizeof(T)); }
    static Mat ones(int r, int c, int type) {
        Mat m(r, c, type);
        for (size_t i = 0; i < m.buf->size(); i += type) (*m.buf)[i] = 1;  // OpenCV: ones() sets channel 0 only
        return m;
    }
};
inline Mat operator*(const Mat& a, int s) {
    Mat m = a.clone();
    for (auto& b : *m.buf) { int v = b * s; b = (uchar)(v > 255 ? 255 : v); }
    return m;
}
}  // namespace cv
namespace cvshim {
// imread() yields a blank image of the PNG's pixel size (so grid_map.cpp:31-32 computes the
// resized W,H itself); resize() then returns the grid the host already resized with cv2.
inline cv::Mat& next_image() { static cv::Mat m; return m; }
inline cv::Size& raw_size() { static cv::Size s; return s; }
inline void set_next_image(const unsigned char* p, int rows, int cols, int raw_rows, int raw_cols) {
    cv::Mat m(rows, cols, 1); memcpy(m.data, p, (size_t)rows * cols); next_image() = m;
    raw_size() = cv::Size(raw_cols, raw_rows);
}
}
namespace cv {
inline Mat imread(const std::string&, int) { return Mat(cvshim::raw_size().height, cvshim::raw_size().width, 1); }
inline void resize(const Mat& src, Mat& dst, Size s) {
    const Mat& g = cvshim::next_image();
    if (src.rows == cvshim::raw_size().height && src.cols == cvshim::raw_size().width && g.rows == s.height && g.cols == s.width) { dst = g.clone(); return; }
    throw std::runtime_error("cv shim: resize() only serves the host-resized occupancy grid");
}
inline void cvtColor(const Mat& src, Mat& dst, int) {
    Mat out(src.rows, src.cols, 3);
    if (src.ch == 1) for (size_t i = 0; i < (size_t)src.rows * src.cols; i++) out.data[3 * i] = out.data[3 * i + 1] = out.data[3 * i + 2] = src.data[i];
    else out = src.clone();
    dst = out;
}
// GUI entry points: only reachable with is_show_gui (always false in the oracle).
inline void line(Mat&, Point, Point, Vec3b, int) {}
inline void circle(Mat&, Point, int, Vec3b) {}
inline void putText(Mat&, const std::string&, Point, int, double, Vec3b, int) {}
inline void rectangle(Mat&, Point, Point, Vec3b, int) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int) { return 0; }
}  // namespace cv
