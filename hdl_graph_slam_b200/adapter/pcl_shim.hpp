// pcl_shim.hpp — TEST-ONLY stand-in for the handful of PCL / Eigen declarations b200_registration.hpp touches, so the
// adapter can be syntax-checked in an image without PCL (tests/test_adapter_syntax.py).  Signatures follow PCL 1.10
// (the distro package of ROS noetic, /root/reference/docker/noetic/Dockerfile:6).  Nothing here is shipped or linked.
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace Eigen {
struct Matrix4f {
  float m[16];
  float* data() { return m; }
  const float* data() const { return m; }
};
}  // namespace Eigen

namespace pcl {
struct alignas(16) PointXYZI {
  float x, y, z, _pad;
  float intensity;
  float _pad2[3];
};
static_assert(sizeof(PointXYZI) == 32, "pcl::PointXYZI is a 32-byte record");

template <typename PointT>
struct PointCloud {
  std::vector<PointT> points;
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
};

namespace search {
template <typename PointT>
class Search {
public:
  using Ptr = std::shared_ptr<Search<PointT>>;
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  using IndicesConstPtr = std::shared_ptr<const std::vector<int>>;
  explicit Search(const std::string& name = "", bool sorted = false) : name_(name), sorted_(sorted) {}
  virtual ~Search() {}
  virtual void setInputCloud(const PointCloudConstPtr&, const IndicesConstPtr& = IndicesConstPtr()) {}
  virtual int nearestKSearch(const PointT&, int, std::vector<int>&, std::vector<float>&) const = 0;
  virtual int radiusSearch(const PointT&, double, std::vector<int>&, std::vector<float>&, unsigned int = 0) const = 0;

protected:
  std::string name_;
  bool sorted_;
};
}  // namespace search

template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
public:
  using Matrix4 = Eigen::Matrix4f;
  using PointCloudSource = PointCloud<PointSource>;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = PointCloud<PointTarget>;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using KdTreePtr = typename search::Search<PointTarget>::Ptr;
  using Ptr = std::shared_ptr<Registration<PointSource, PointTarget, Scalar>>;
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; target_cloud_updated_ = true; }
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) { tree_ = tree; force_no_recompute_ = force_no_recompute; }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  void align(PointCloudSource& output, const Matrix4& guess) {
    output.points = input_->points;
    converged_ = false;
    computeTransformation(output, guess);
  }

protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  KdTreePtr tree_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0, max_iterations_ = 10;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_ = 0, corr_dist_threshold_ = 0;
  bool converged_ = false, target_cloud_updated_ = true, force_no_recompute_ = false;
};
// stand-in for pcl::transformPointCloud (pcl/common/transforms.h); column-major 4x4 like Eigen
template <typename PointT>
inline void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  out.points = in.points;
  for (auto& p : out.points) {
    const float x = p.x, y = p.y, z = p.z;
    p.x = T.m[0] * x + T.m[4] * y + T.m[8] * z + T.m[12];
    p.y = T.m[1] * x + T.m[5] * y + T.m[9] * z + T.m[13];
    p.z = T.m[2] * x + T.m[6] * y + T.m[10] * z + T.m[14];
  }
}
}  // namespace pcl
