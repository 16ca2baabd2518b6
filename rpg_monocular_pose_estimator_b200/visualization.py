"""Host mirror of monocular_pose_estimator::Visualization (monocular_pose_estimator_lib/src/visualization.cpp:37-104): the debug
overlay MPENode publishes when someone subscribes to it (monocular_pose_estimator.cpp:198-214).  Like the reference it runs on
the host with OpenCV's drawing functions — it is not on the pose path and nothing of it runs on the GPU; the pose, the region of
interest and the detection centres it draws are the ones the device returned in the result record.  Pinned against the
reference's own source in tests/test_oracle_pose_ref.py."""
from __future__ import annotations

import numpy as np


class Visualization:
    @staticmethod
    def projectOrientationVectorsOnImage(image, points_to_project, camera_matrix_K, camera_distortion_coeffs):
        """visualization.cpp:37-55: the three axes of the body frame, red / green / blue, 2 px wide (colours are CV_RGB, i.e. the
        image is taken to be BGR)."""
        import cv2
        pts = np.asarray(points_to_project, np.float32).reshape(-1, 1, 3)           # std::vector<cv::Point3f>
        proj, _ = cv2.projectPoints(pts, np.zeros((3, 1)), np.zeros((3, 1)), np.asarray(camera_matrix_K, np.float64),
                                    np.asarray(camera_distortion_coeffs, np.float64))
        p = proj.reshape(-1, 2).astype(np.float32)                                  # std::vector<cv::Point2f>
        as_point = lambda q: (int(np.rint(q[0])), int(np.rint(q[1])))               # cv::Point(cv::Point2f): saturate_cast<int> = cvRound
        cv2.line(image, as_point(p[0]), as_point(p[1]), (0, 0, 255, 0), 2)
        cv2.line(image, as_point(p[0]), as_point(p[2]), (0, 255, 0, 0), 2)
        cv2.line(image, as_point(p[0]), as_point(p[3]), (255, 0, 0, 0), 2)

    @staticmethod
    def createVisualizationImage(image, transform, camera_matrix_K, camera_distortion_coeffs, region_of_interest,
                                 distorted_detection_centers):
        """visualization.cpp:57-104, in place on a 3-channel 8-bit image."""
        import cv2
        length = 0.075                                                               # orientation_vector_length
        T = np.asarray(transform, np.float64).reshape(4, 4)
        O = np.array([[0, length, 0, 0], [0, 0, length, 0], [0, 0, 0, length], [1, 1, 1, 1]], np.float64)
        vis = np.zeros((4, 4))
        for r in range(4):                                                           # transform * orientation_vector_points, k ascending
            for c in range(4):
                acc = T[r, 0] * O[0, c]
                for k in range(1, 4):
                    acc = acc + T[r, k] * O[k, c]
                vis[r, c] = acc
        pts = vis[:3, :].T.astype(np.float32)                                        # cv::Point3f(x, y, z) of columns 0..3
        Visualization.projectOrientationVectorsOnImage(image, pts, camera_matrix_K, camera_distortion_coeffs)
        for c in np.asarray(distorted_detection_centers, np.float32).reshape(-1, 2):
            cv2.circle(image, (int(np.rint(c[0])), int(np.rint(c[1]))), 10, (0, 0, 255, 0), 2)
        x, y, w, h = [int(v) for v in region_of_interest]
        cv2.rectangle(image, (x, y, w, h), (255, 0, 0, 0), 2)
        return image
