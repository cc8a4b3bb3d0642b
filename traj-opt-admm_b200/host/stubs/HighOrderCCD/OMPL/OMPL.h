// Stand-in for HighOrderCCD/OMPL/OMPL.h (reference :17-52): the RRT-Connect front end needs the system OMPL library,
// which is neither vendored by the reference nor installed here, and is only used when "init":2.  Same class name,
// constructor and the two members Main calls; planRRT reports failure so the caller keeps its own way-points.
#ifndef OMPL_H
#define OMPL_H

#include "HighOrderCCD/Utils/CCDUtils.h"
#include "HighOrderCCD/CCD/CCD.h"
#include "HighOrderCCD/BVH/BVH.h"

#include <vector>

PRJ_BEGIN

class OMPL {
 public:
  OMPL(Eigen::VectorXd, Eigen::VectorXd, Eigen::MatrixXd, std::vector<std::vector<Eigen::MatrixXd>>, BVH&) {}
  int nrBroad() const { return 0; }
  int nrNarrow() const { return 0; }
  void getPath(std::vector<Eigen::Vector3d>& path) { path = _path; }
  bool planRRT(Eigen::Vector3d, Eigen::Vector3d, Eigen::MatrixXd, std::vector<std::vector<Eigen::MatrixXd>>, BVH&, int = 1200) {
    std::cerr << "OMPL: built without the OMPL library; use \"init\":1 (way-point file)" << std::endl;
    return false;
  }

 protected:
  std::vector<Eigen::Vector3d> _path;
};

PRJ_END

#endif
