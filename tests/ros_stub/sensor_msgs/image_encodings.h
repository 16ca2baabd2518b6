// TEST INFRASTRUCTURE — stand-in for <sensor_msgs/image_encodings.h>
#pragma once
#include <string>
namespace sensor_msgs { namespace image_encodings {
const std::string MONO8 = "mono8";
const std::string BGR8 = "bgr8";
} }
