// TEST INFRASTRUCTURE — stand-in for <sensor_msgs/Image.h> (fields of the public message definition)
#pragma once
#include "ros/ros.h"
namespace sensor_msgs {
struct Image {
  std_msgs::Header header;
  uint32_t height, width;
  std::string encoding;
  uint8_t is_bigendian;
  uint32_t step;
  std::vector<uint8_t> data;
  Image() : height(0), width(0), is_bigendian(0), step(0) {}
  typedef boost::shared_ptr<Image> Ptr;
  typedef boost::shared_ptr<Image const> ConstPtr;
};
typedef Image::Ptr ImagePtr;
typedef Image::ConstPtr ImageConstPtr;
}  // namespace sensor_msgs
