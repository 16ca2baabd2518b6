// TEST INFRASTRUCTURE — stand-in for <pluginlib/class_list_macros.h>: the export macro becomes a factory function the harness can call
#pragma once
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type) \
  extern "C" base_class_type* ros_stub_create_plugin() { return new class_type(); }
