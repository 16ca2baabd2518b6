// TEST INFRASTRUCTURE — stand-in for <sensor_msgs/CameraInfo.h>
#pragma once
#include "ros/ros.h"
namespace sensor_msgs {
struct CameraInfo {
  std_msgs::Header header;
  uint32_t height, width;
  std::string distortion_model;
  std::vector<double> D;
  boost::array<double, 9> K;
  boost::array<double, 9> R;
  boost::array<double, 12> P;
  CameraInfo() : height(0), width(0), K(), R(), P() {}
  typedef boost::shared_ptr<CameraInfo> Ptr;
  typedef boost::shared_ptr<CameraInfo const> ConstPtr;
};
}  // namespace sensor_msgs
