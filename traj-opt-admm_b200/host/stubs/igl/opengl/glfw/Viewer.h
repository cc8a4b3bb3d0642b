// Headless stand-in for libigl's igl/opengl/glfw/Viewer.h: Main/admmPathPlanning3D.cpp and Main/multiPathPlanning3D.cpp
// include it unconditionally but only touch it inside their `gui` branch ("gui":0 in Config File/3D.json).  The real
// header needs GLFW/glad, which libigl downloads at configure time (lib/libigl/cmake/libigl.cmake:360-387).
// Only the members those two files use exist; launch() reports that there is no display and returns.
#ifndef TRAJOPT_HEADLESS_VIEWER_H
#define TRAJOPT_HEADLESS_VIEWER_H

#include <Eigen/Core>
#include <functional>
#include <iostream>

namespace igl { namespace opengl {

struct ViewerCore {
  Eigen::Vector4f background_color = Eigen::Vector4f(1, 1, 1, 1);
  bool is_animating = false;
  float camera_zoom = 1.0f;
};

struct ViewerData {
  float line_width = 1.0f, point_size = 1.0f;
  template <typename A, typename B> void set_points(const A&, const B&) {}
  template <typename A, typename B> void add_points(const A&, const B&) {}
  template <typename A, typename B, typename C> void add_edges(const A&, const B&, const C&) {}
  void clear_edges() {}
};

namespace glfw {

class Viewer {
 public:
  ViewerCore& core() { return core_; }
  ViewerData& data() { return data_; }
  std::function<bool(Viewer&)> callback_pre_draw;
  std::function<bool(Viewer&, unsigned char, int)> callback_key_down;
  int launch() {
    std::cerr << "igl::opengl::glfw::Viewer: headless build, no window (set \"gui\":0)" << std::endl;
    return 1;
  }

 private:
  ViewerCore core_;
  ViewerData data_;
};

}}}  // namespace igl::opengl::glfw
#endif
