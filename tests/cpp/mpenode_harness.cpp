// Test driver for the reference's ROS node, UNMODIFIED (monocular_pose_estimator/src/monocular_pose_estimator.cpp, compiled from
// where it lies), on top of the class shim of this repository: the node's `#include "monocular_pose_estimator_lib/pose_estimator.h"`
// resolves to tests/ros_stub/monocular_pose_estimator_lib/pose_estimator.h -> include/monocular_pose_estimator_b200/shim.h, and
// ROS / cv_bridge / dynamic_reconfigure are the in-process stand-ins of tests/ros_stub (Eigen / OpenCV: the stand-ins of oracle/).
// The harness plays roscore: marker list on the parameter server, one CameraInfo, a reconfigure call, then one sensor_msgs/Image
// per frame of a scene file (same format as tests/cpp/shim_real_types_demo.cpp), and prints what the node published.  With the
// argument `nodelet` the node is created the way the nodelet manager does it (the reference's nodelet.cpp, also unmodified).
//   per frame:  index  published(0/1)  roi x y w h  gauss-newton iterations  pose 4x4 row-major  |  position xyz  orientation xyzw  cov[0]
// everything the node header includes, first (all guarded), so that the access trick below touches the MPENode class only
#include <sstream>
#include "ros/ros.h"
#include <sensor_msgs/Image.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/image_encodings.h>
#include <geometry_msgs/PoseWithCovarianceStamped.h>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>
#include <image_transport/image_transport.h>
#include <cv_bridge/cv_bridge.h>
#include <dynamic_reconfigure/server.h>
#include <monocular_pose_estimator/MonocularPoseEstimatorConfig.h>
#include "monocular_pose_estimator_lib/pose_estimator.h"
#include <nodelet/nodelet.h>
#define private public            // the harness reads MPENode::trackable_object_ (ROI, iteration count); the node's own TU is untouched
#define protected public          // ... and MPENodelet::mpe_node
#include "monocular_pose_estimator/monocular_pose_estimator.h"
#include "monocular_pose_estimator/nodelet.h"
#undef protected
#undef private

extern "C" nodelet::Nodelet* ros_stub_create_plugin();   // what PLUGINLIB_EXPORT_CLASS in the reference's nodelet.cpp expands to here

#include <cstdio>
#include <cstring>

using namespace monocular_pose_estimator;

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: mpenode_on_shim scene.bin [nodelet]\n"); return 2; }
  const bool as_nodelet = argc > 2 && std::strcmp(argv[2], "nodelet") == 0;   // load the node the way the nodelet manager does
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int hdr[4];
  if (fread(hdr, sizeof(int), 4, f) != 4) return 2;
  const int n_frames = hdr[0], w = hdr[1], h = hdr[2], n_leds = hdr[3];
  double K[9], D[5], params[11];
  std::vector<double> markers(3 * n_leds), times(n_frames);
  if (fread(K, 8, 9, f) != 9 || fread(D, 8, 5, f) != 5 || fread(markers.data(), 8, markers.size(), f) != markers.size() ||
      fread(params, 8, 11, f) != 11 || fread(times.data(), 8, times.size(), f) != times.size()) return 2;
  std::vector<uint8_t> frames((size_t)n_frames * w * h);
  if (fread(frames.data(), 1, frames.size(), f) != frames.size()) return 2;
  fclose(f);

  try {
    ros_stub::Bus& bus = ros_stub::bus();
    // rosparam: ~marker_positions = [{x, y, z}, ...]   (monocular_pose_estimator.cpp:62-82)
    XmlRpc::XmlRpcValue list;
    for (int i = 0; i < n_leds; ++i) { list[i]["x"] = markers[3 * i]; list[i]["y"] = markers[3 * i + 1]; list[i]["z"] = markers[3 * i + 2]; }
    bus.params["marker_positions"] = list;
    bool got_pose = false;
    geometry_msgs::PoseWithCovarianceStamped last_pose;
    size_t overlay_bytes = 0;
    bus.sinks["estimated_pose"] = [&](const void* m) { last_pose = *static_cast<const geometry_msgs::PoseWithCovarianceStamped*>(m); got_pose = true; };
    bus.sinks["image_with_detections"] = [&](const void* m) { overlay_bytes = static_cast<const sensor_msgs::Image*>(m)->data.size(); };

    std::unique_ptr<MPENode> direct;
    std::unique_ptr<nodelet::Nodelet> plugin;
    MPENode* node_ptr = nullptr;
    if (as_nodelet) {
      plugin.reset(ros_stub_create_plugin());                       // PLUGINLIB_EXPORT_CLASS(monocular_pose_estimator::MPENodelet, nodelet::Nodelet)
      plugin->init();                                               // MPENodelet::onInit (nodelet.cpp:24-28)
      node_ptr = static_cast<MPENodelet*>(plugin.get())->mpe_node.get();
    } else {
      direct.reset(new MPENode((ros::NodeHandle()), ros::NodeHandle("~")));   // node.cpp:27
      node_ptr = direct.get();
    }
    MPENode& node = *node_ptr;
    if (bus.shutdown_requested) { fprintf(stderr, "node asked for shutdown\n"); return 1; }

    MonocularPoseEstimatorConfig cfg;                               // dynamic_reconfigure pushes the launch file's values
    cfg.threshold_value = (int)params[0]; cfg.gaussian_sigma = params[1]; cfg.min_blob_area = params[2]; cfg.max_blob_area = params[3];
    cfg.max_width_height_distortion = params[4]; cfg.max_circular_distortion = params[5]; cfg.back_projection_pixel_tolerance = params[6];
    cfg.nearest_neighbour_pixel_tolerance = params[7]; cfg.certainty_threshold = params[8]; cfg.valid_correspondence_threshold = params[9];
    cfg.roi_border_thickness = (int)params[10];
    ros_stub::reconfigure<MonocularPoseEstimatorConfig>()(cfg, 0);

    sensor_msgs::CameraInfo::Ptr info(new sensor_msgs::CameraInfo());
    info->width = (uint32_t)w; info->height = (uint32_t)h; info->D.assign(D, D + 5);
    for (int i = 0; i < 9; ++i) info->K[i] = K[i];
    sensor_msgs::CameraInfo::ConstPtr cinfo = info;
    bus.subscribers["/camera/camera_info"](&cinfo);

    for (int fi = 0; fi < n_frames; ++fi) {
      sensor_msgs::Image::Ptr img(new sensor_msgs::Image());
      img->header.stamp = ros::Time(times[fi]); img->header.seq = (uint32_t)fi;
      img->width = (uint32_t)w; img->height = (uint32_t)h; img->step = (uint32_t)w; img->encoding = "mono8";
      img->data.assign(frames.begin() + (size_t)fi * w * h, frames.begin() + (size_t)(fi + 1) * w * h);
      sensor_msgs::Image::ConstPtr cimg = img;
      bus.num_subscribers["image_with_detections"] = (fi % 4 == 0) ? 1u : 0u;        // somebody watches every fourth image
      got_pose = false; overlay_bytes = 0;
      bus.subscribers["/camera/image_raw"](&cimg);
      if (fi % 4 == 0 && overlay_bytes != (size_t)3 * w * h) { fprintf(stderr, "no BGR overlay image published for frame %d\n", fi); return 1; }
      const cv::Rect& r = node.trackable_object_.regionOfInterest();
      Eigen::Matrix4d T = node.trackable_object_.getPredictedPose();
      printf("%d %d %d %d %d %d %u", fi, got_pose ? 1 : 0, r.x, r.y, r.width, r.height, node.trackable_object_.lastGaussNewtonIterations());
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) printf(" %.17g", T(a, b));
      printf(" | %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", last_pose.pose.pose.position.x, last_pose.pose.pose.position.y,
             last_pose.pose.pose.position.z, last_pose.pose.pose.orientation.x, last_pose.pose.pose.orientation.y, last_pose.pose.pose.orientation.z,
             last_pose.pose.pose.orientation.w, last_pose.pose.covariance.elems[0]);
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "mpenode_on_shim: %s\n", e.what());
    return 1;
  }
  return 0;
}
