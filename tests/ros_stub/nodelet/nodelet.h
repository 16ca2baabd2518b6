// TEST INFRASTRUCTURE — stand-in for <nodelet/nodelet.h>: the base class MPENodelet derives from (nodelet.h:8-14, nodelet.cpp:24-28)
#pragma once
#include "ros/ros.h"
namespace nodelet {
class Nodelet {
 public:
  virtual ~Nodelet() {}
  virtual void onInit() = 0;
  void init() { onInit(); }                                  // what the nodelet manager calls after loading the plugin
 protected:
  ros::NodeHandle& getNodeHandle() { return nh_; }
  ros::NodeHandle& getPrivateNodeHandle() { return private_nh_; }
  const std::string& getName() const { return name_; }
 private:
  ros::NodeHandle nh_;
  ros::NodeHandle private_nh_{"~"};
  std::string name_ = "mpe_stub_nodelet";
};
}  // namespace nodelet
#define NODELET_INFO_STREAM(x) do { } while (0)
