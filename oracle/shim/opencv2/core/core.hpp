// Minimal stand-in for <opencv2/core/core.hpp>, used ONLY to compile the reference's
// pairwise3d.h (which needs cv::Point2f / cv::Point3f in two factory signatures,
// pairwise3d.h:5,38,52) as the parity oracle.  Test infrastructure, not product code.
#pragma once
namespace cv {
struct Point2f {
    float x, y;
    Point2f(float x_ = 0, float y_ = 0) : x(x_), y(y_) {}
};
struct Point3f {
    float x, y, z;
    Point3f(float x_ = 0, float y_ = 0, float z_ = 0) : x(x_), y(y_), z(z_) {}
};
}  // namespace cv
