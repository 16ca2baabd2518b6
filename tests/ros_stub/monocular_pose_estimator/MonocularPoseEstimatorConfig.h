// TEST INFRASTRUCTURE — what dynamic_reconfigure generates from monocular_pose_estimator/cfg/MonocularPoseEstimator.cfg:12-22
// (names, types and defaults of the eleven parameters)
#pragma once
namespace monocular_pose_estimator {
struct MonocularPoseEstimatorConfig {
  int threshold_value;
  double gaussian_sigma, min_blob_area, max_blob_area, max_width_height_distortion, max_circular_distortion;
  double back_projection_pixel_tolerance, nearest_neighbour_pixel_tolerance, certainty_threshold, valid_correspondence_threshold;
  int roi_border_thickness;
  MonocularPoseEstimatorConfig()
      : threshold_value(180), gaussian_sigma(0.6), min_blob_area(10), max_blob_area(200), max_width_height_distortion(0.5),
        max_circular_distortion(0.5), back_projection_pixel_tolerance(5), nearest_neighbour_pixel_tolerance(5), certainty_threshold(0.75),
        valid_correspondence_threshold(0.7), roi_border_thickness(10) {}
};
}  // namespace monocular_pose_estimator
