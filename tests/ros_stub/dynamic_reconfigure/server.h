// TEST INFRASTRUCTURE — stand-in for <dynamic_reconfigure/server.h>: setCallback calls back at once with the defaults, as the real
// server does; the harness pushes further configurations through ros_stub::reconfigure<Config>()
#pragma once
#include "ros/ros.h"
namespace ros_stub {
template <typename C> std::function<void(C&, uint32_t)>& reconfigure() { static std::function<void(C&, uint32_t)> cb; return cb; }
}
namespace dynamic_reconfigure {
template <typename ConfigType> class Server {
 public:
  typedef boost::function<void(ConfigType&, uint32_t)> CallbackType;
  void setCallback(const CallbackType& cb) {
    ros_stub::reconfigure<ConfigType>() = cb;
    ConfigType defaults;
    cb(defaults, ~0u);
  }
};
}  // namespace dynamic_reconfigure
