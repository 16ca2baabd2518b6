"""TEST INFRASTRUCTURE — CPU oracle of the reference hot path (see oracle/README in DESIGN.md §oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (rpg_monocular_pose_estimator_b200) never does.
"""
