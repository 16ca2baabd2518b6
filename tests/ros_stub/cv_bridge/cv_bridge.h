// TEST INFRASTRUCTURE — stand-in for <cv_bridge/cv_bridge.h>: MONO8 in, MONO8 / BGR8 out (what MPENode uses)
#pragma once
#include <cstring>
#include <stdexcept>
#include <opencv2/opencv.hpp>
#include "sensor_msgs/Image.h"
namespace cv_bridge {
class Exception : public std::runtime_error { public: explicit Exception(const std::string& w) : std::runtime_error(w) {} };
class CvImage {
 public:
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  sensor_msgs::ImagePtr toImageMsg() const {
    sensor_msgs::ImagePtr m(new sensor_msgs::Image());
    const int ch = image.channels();
    m->header = header; m->encoding = encoding; m->height = (uint32_t)image.rows; m->width = (uint32_t)image.cols; m->step = (uint32_t)(image.cols * ch);
    m->data.resize((std::size_t)m->step * m->height);
    for (int r = 0; r < image.rows; ++r) std::memcpy(&m->data[(std::size_t)r * m->step], image.data + (std::size_t)r * image.step, m->step);
    return m;
  }
};
typedef boost::shared_ptr<CvImage> CvImagePtr;
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& src, const std::string& encoding) {
  if (src->encoding != "mono8" || encoding != "mono8") throw Exception("stand-in cv_bridge converts mono8 to mono8 only");
  CvImagePtr p(new CvImage());
  p->header = src->header; p->encoding = encoding;
  p->image = cv::Mat((int)src->height, (int)src->width, CV_8UC1);
  for (uint32_t r = 0; r < src->height; ++r) std::memcpy(p->image.data + (std::size_t)r * p->image.step, &src->data[(std::size_t)r * src->step], src->width);
  return p;
}
}  // namespace cv_bridge
