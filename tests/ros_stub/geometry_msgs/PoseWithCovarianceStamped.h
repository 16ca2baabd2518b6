// TEST INFRASTRUCTURE — stand-in for <geometry_msgs/PoseWithCovarianceStamped.h>
#pragma once
#include "ros/ros.h"
namespace geometry_msgs {
struct Point { double x, y, z; Point() : x(0), y(0), z(0) {} };
struct Quaternion { double x, y, z, w; Quaternion() : x(0), y(0), z(0), w(0) {} };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; boost::array<double, 36> covariance; PoseWithCovariance() : covariance() {} };
struct PoseWithCovarianceStamped { std_msgs::Header header; PoseWithCovariance pose; };
}  // namespace geometry_msgs
