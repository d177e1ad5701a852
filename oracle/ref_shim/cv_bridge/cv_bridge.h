// TEST INFRASTRUCTURE shim: cv_bridge::CvImage::toImageMsg is a byte copy of the Mat.
#pragma once
#include <opencv2/opencv.hpp>
#include <sensor_msgs/Image.h>
namespace cv_bridge {
struct CvImage {
    std_msgs::Header header; std::string encoding; cv::Mat image;
    CvImage() {}
    CvImage(const std_msgs::Header& h, const std::string& e, const cv::Mat& m) : header(h), encoding(e), image(m) {}
    void toImageMsg(sensor_msgs::Image& out) const {
        out.header = header; out.encoding = encoding; out.height = image.rows; out.width = image.cols;
        out.step = image.cols * image.ch;
        out.data.assign(image.data, image.data + (size_t)image.rows * image.cols * image.ch);
    }
};
}
