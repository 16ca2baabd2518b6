// TEST INFRASTRUCTURE — stand-in for <image_transport/image_transport.h>
#pragma once
#include "ros/ros.h"
#include "sensor_msgs/Image.h"
namespace image_transport {
class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& topic) : topic_(topic) {}
  uint32_t getNumSubscribers() const { return ros_stub::bus().num_subscribers[topic_]; }
  void publish(const sensor_msgs::ImagePtr& m) const {
    auto it = ros_stub::bus().sinks.find(topic_);
    if (it != ros_stub::bus().sinks.end()) it->second(m.get());
  }
 private:
  std::string topic_;
};
class ImageTransport {
 public:
  explicit ImageTransport(const ros::NodeHandle&) {}
  Publisher advertise(const std::string& topic, uint32_t /*queue*/) { return Publisher(topic); }
};
}  // namespace image_transport
