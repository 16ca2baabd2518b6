// TEST INFRASTRUCTURE — a stand-in for <ros/ros.h>, just large enough to compile the reference's ROS node
// (monocular_pose_estimator/src/monocular_pose_estimator.cpp) UNMODIFIED against the C++ shim of this repository and to drive its
// callbacks from a test harness (tests/cpp/mpenode_harness.cpp).  Not ROS code: a tiny in-process "bus" with the handful of
// calls MPENode makes (monocular_pose_estimator.cpp:39-236), written from the public API.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

// ---- the parts of boost the node reaches through the ROS headers -----------------------------------------------------------
namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename S> using function = std::function<S>;
using std::bind;
template <typename T, std::size_t N> struct array {          // boost::array exposes `elems`
  T elems[N];
  T& operator[](std::size_t i) { return elems[i]; }
  const T& operator[](std::size_t i) const { return elems[i]; }
  static std::size_t size() { return N; }
};
}  // namespace boost
using namespace std::placeholders;                            // <boost/bind.hpp> puts _1, _2 into the global namespace

// ---- XmlRpc::XmlRpcValue: the marker list arrives as an array of {x, y, z} structs (monocular_pose_estimator.cpp:62-82) --------
namespace XmlRpc {
class XmlRpcValue {
 public:
  XmlRpcValue() : d_(0) {}
  XmlRpcValue(double d) : d_(d) {}
  int size() const { return (int)items_.size(); }
  XmlRpcValue& operator[](int i) { if ((int)items_.size() <= i) items_.resize((std::size_t)i + 1); return items_[(std::size_t)i]; }
  XmlRpcValue& operator[](const char* k) { return members_[k]; }
  operator double&() { return d_; }
 private:
  double d_;
  std::vector<XmlRpcValue> items_;
  std::map<std::string, XmlRpcValue> members_;
};
}  // namespace XmlRpc

namespace ros_stub {
// everything the node registers or publishes, type-erased; the harness looks callbacks up by topic
struct Bus {
  std::map<std::string, std::function<void(const void*)> > subscribers;   // argument: const boost::shared_ptr<M const>*
  std::map<std::string, std::function<void(const void*)> > sinks;         // argument: const M* (what a publisher sent)
  std::map<std::string, XmlRpc::XmlRpcValue> params;
  std::map<std::string, unsigned> num_subscribers;                          // of the topics the node publishes
  bool shutdown_requested = false;
};
inline Bus& bus() { static Bus b; return b; }
}  // namespace ros_stub

namespace ros {
struct Time {
  double sec_;
  Time(double s = 0) : sec_(s) {}
  double toSec() const { return sec_; }
};
class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& topic) : topic_(topic) {}
  template <typename M> void publish(const M& m) const {
    auto it = ros_stub::bus().sinks.find(topic_);
    if (it != ros_stub::bus().sinks.end()) it->second(&m);
  }
  uint32_t getNumSubscribers() const { return ros_stub::bus().num_subscribers[topic_]; }
 private:
  std::string topic_;
};
class Subscriber {};
class NodeHandle {
 public:
  NodeHandle() {}
  NodeHandle(const std::string& ns) : ns_(ns) {}
  template <typename M, typename T>
  Subscriber subscribe(const std::string& topic, uint32_t /*queue*/, void (T::*fp)(const boost::shared_ptr<M const>&), T* obj) {
    ros_stub::bus().subscribers[topic] = [fp, obj](const void* msg) { (obj->*fp)(*static_cast<const boost::shared_ptr<M const>*>(msg)); };
    return Subscriber();
  }
  template <typename M> Publisher advertise(const std::string& topic, uint32_t /*queue*/) { return Publisher(topic); }
  bool getParam(const std::string& key, XmlRpc::XmlRpcValue& v) const {
    auto it = ros_stub::bus().params.find(key);
    if (it == ros_stub::bus().params.end()) return false;
    v = it->second;
    return true;
  }
 private:
  std::string ns_;
};
namespace this_node { inline std::string getName() { return "mpe_stub_node"; } }
inline void shutdown() { ros_stub::bus().shutdown_requested = true; }
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
}  // namespace ros

#define ROS_STUB_LOG(tag, ...) do { std::fprintf(stderr, "[" tag "] "); std::fprintf(stderr, __VA_ARGS__); std::fputc('\n', stderr); } while (0)
#define ROS_ERROR(...) ROS_STUB_LOG("ERROR", __VA_ARGS__)
#define ROS_WARN(...) do { } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_DEBUG_STREAM(x) do { } while (0)

namespace std_msgs {
struct Header { uint32_t seq; ros::Time stamp; std::string frame_id; Header() : seq(0) {} };
}  // namespace std_msgs
