// opencv2/ml/ml.hpp — STAND-IN (test infrastructure only) for cv::ml::EM, the third-party mixture fit of the reference's
// GMM thread (particle_filter.cpp:252-318).  It is NOT OpenCV's algorithm: every fit returns ONE cluster, the sample
// mean and the maximum-likelihood covariance of the sample matrix, and likelihoods of zero (so computeGMM never changes
// its cluster count).  That is enough to drive the code AROUND the fit — the sample matrix (:262-272) and the adaptive
// particle count update() derives from the covariances (:151-158) — which is what this build pins.
#pragma once
#include <opencv2/core/core.hpp>
namespace cv { namespace ml {
class EM {
 public:
  enum { COV_MAT_GENERIC = 2 };
  static Ptr<EM> create() { return std::make_shared<EM>(); }
  void setCovarianceMatrixType(int) {}
  void setClustersNumber(int) {}
  bool trainEM(const Mat& samples, Mat& likelihoods, Mat& labels) {
    last_samples_ = Mat(samples.rows, samples.cols, CV_64F);   // a copy for the harness: the matrix computeGMM built
    std::memcpy(last_samples_.data, samples.data, samples.total() * 8);
    const int n = samples.rows, k = samples.cols;
    means_ = Mat(1, k, CV_64F); cov_ = Mat(k, k, CV_64F);
    for (int i = 0; i < n; i++) for (int j = 0; j < k; j++) means_.at<double>(0, j) += samples.at<double>(i, j) / n;
    for (int i = 0; i < n; i++) for (int a = 0; a < k; a++) for (int b = 0; b < k; b++)
      cov_.at<double>(a, b) += (samples.at<double>(i, a) - means_.at<double>(0, a)) * (samples.at<double>(i, b) - means_.at<double>(0, b)) / n;
    likelihoods = Mat(n, 1, CV_64F); labels = Mat(n, 1, CV_64F);
    return true;
  }
  Mat getMeans() const { return means_; }
  void getCovs(std::vector<Mat>& covs) const { covs.assign(1, cov_); }
  static Mat& lastSamples() { return last_samples_; }
 private:
  Mat means_, cov_;
  static inline thread_local Mat last_samples_;   // per thread: the GMM threads of live filters keep fitting
};
}}  // namespace cv::ml
