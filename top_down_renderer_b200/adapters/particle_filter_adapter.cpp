// adapters/particle_filter_adapter.cpp — the bodies that REPLACE src/particle_filter.cpp and src/state_particle.cpp of the reference.
// The class declarations are the reference's own, unchanged (include/top_down_render/particle_filter.h:22-73,
// state_particle.h:9-71).  Scoring, normalisation, resampling, propagation and the pose run on the device through the C ABI
// (include/tdr.h); what stays host code is what is sequential on the ONE shared std::mt19937 in the reference too — the
// rejection sampling of the StateParticle constructor, the standard-normal variates of propagate, the single uniform of
// update — so a run with the same seed consumes the engine exactly as the reference does.
//
// particles_ / new_particles_ / weights_ — the members the header declares — are kept as host mirrors of the device set
// (states, last_dist, raw weight per particle; normalised weights), which is what visualize, the GMM thread and the harness
// read.  The headers give ParticleFilter no access to StateParticle's private fields, so the reference's own call structure
// carries the results back: update() scores the WHOLE set on the device, then calls computeWeight on every particle, whose
// body here takes that particle's raw weight / heading from the staged result; propagate() likewise.  A production build
// would refresh lazily (28 B per particle and step over PCIe); tdr_host.hpp shows that variant.  visualize (:367-421,
// drawing only) keeps the reference's body.
#include <cstring>
#include <vector>

#include "top_down_render/particle_filter.h"

#include "tdr_adapter_common.h"

using tdr_adapter::ok;
using tdr_adapter::tdr;
static_assert(sizeof(State) == sizeof(tdr_state), "State and tdr_state share their layout");

// ============================ StateParticle (state_particle.cpp) ============================================================
// :3-49 — host code: every draw depends on where the previous rejection left the shared engine
StateParticle::StateParticle(std::mt19937* gen, TopDownMapPolar* map, FilterParams* params, bool init) {
  params_ = params;
  gen_ = gen;
  last_dist_ = 0;
  std::uniform_real_distribution<float> uniform_dist(0., 1.);
  std::normal_distribution<float> normal_dist(0., 1.);
  Eigen::Vector2f map_size = map->size().cast<float>() * map->resolution();
  if (init) {
    state_.scale = params_->fixed_scale < 0 ? std::pow(10, (uniform_dist(*gen) - 0.5) * 2) : params_->fixed_scale;
    std::vector<int> here;
    do {
      if (params_->init_pos_px_x > 0) {
        state_.init_x_px = std::clamp<float>(normal_dist(*gen) * params_->init_pos_px_cov + params_->init_pos_px_x, 0, map_size[0]);
        state_.init_y_px = std::clamp<float>(normal_dist(*gen) * params_->init_pos_px_cov + params_->init_pos_px_y, 0, map_size[1]);
      } else {
        state_.init_x_px = uniform_dist(*gen) * map_size[0];
        state_.init_y_px = uniform_dist(*gen) * map_size[1];
      }
      map->getClassesAtPoint(Eigen::Vector2i(state_.init_x_px, state_.init_y_px), here);
    } while (std::find(here.begin(), here.end(), 1) == here.end());          // until the particle sits on the road
    if (params_->init_pos_deg_theta != std::numeric_limits<float>::infinity()) {
      state_.theta = normal_dist(*gen) * params_->init_pos_deg_cov + params_->init_pos_deg_theta;
      state_.theta *= M_PI / 180;
      state_.have_init = true;
    } else {
      state_.theta = 0;
      state_.have_init = false;
    }
  }
  map_ = map;
  width_ = map_size[0];
  height_ = map_size[1];
  weight_ = 0;
}
void StateParticle::updateSize() {
  Eigen::Vector2f map_size = map_->size().cast<float>() * map_->resolution();
  width_ = map_size[0];
  height_ = map_size[1];
}
void StateParticle::setState(const State& s) {
  state_.init_x_px = s.init_x_px; state_.init_y_px = s.init_y_px; state_.dx_m = s.dx_m; state_.dy_m = s.dy_m;
  state_.theta = s.theta; state_.scale = s.scale; state_.have_init = s.have_init;
}
void StateParticle::setScale(float scale) { state_.scale = scale; }
State StateParticle::state() const { return state_; }
Eigen::Vector4f StateParticle::mlState() {                                    // :98-102
  return Eigen::Vector4f(state_.dx_m * state_.scale + state_.init_x_px, state_.dy_m * state_.scale + state_.init_y_px, state_.theta, state_.scale);
}
float StateParticle::weight() const { return weight_; }
float StateParticle::lastDist() const { return last_dist_; }

// propagate (:57-78) and computeWeight / getCostForRot (:112-219) run for the WHOLE set on the device (ParticleFilter below);
// these two members hand one particle its share of the staged result, in particle order
namespace {
std::vector<tdr_state> g_stage_states;
std::vector<float> g_stage_last_dist, g_stage_weight;
size_t g_cursor = 0;
}  // namespace
void StateParticle::propagate(Eigen::Vector2f&, float, bool) {
  const tdr_state& t = g_stage_states[g_cursor];
  state_.dx_m = t.dx_m; state_.dy_m = t.dy_m; state_.theta = t.theta; state_.scale = t.scale;
  last_dist_ = g_stage_last_dist[g_cursor++];
}
void StateParticle::computeWeight(std::vector<Eigen::ArrayXXf>&, std::vector<Eigen::ArrayXXf>&, float) {
  const tdr_state& t = g_stage_states[g_cursor];
  state_.theta = t.theta;                                                     // the heading search's choice (:195-206)
  state_.have_init = t.have_init != 0;
  weight_ = g_stage_weight[g_cursor++];
}

// ============================ ParticleFilter (particle_filter.cpp) ==========================================================
namespace {
// the device holds one particle set: the filter that used it last
const void* g_owner = nullptr;

std::vector<tdr_state> states_of(const std::vector<std::shared_ptr<StateParticle>>& v, std::vector<float>& last_dist) {
  std::vector<tdr_state> st(v.size());
  last_dist.resize(v.size());
  for (size_t i = 0; i < v.size(); i++) {
    const State s = v[i]->state();
    std::memset(&st[i], 0, sizeof(tdr_state));
    st[i].init_x_px = s.init_x_px; st[i].init_y_px = s.init_y_px; st[i].dx_m = s.dx_m; st[i].dy_m = s.dy_m;
    st[i].theta = s.theta; st[i].scale = s.scale; st[i].have_init = s.have_init ? 1 : 0;
    last_dist[i] = v[i]->lastDist();
  }
  return st;
}
}  // namespace

// :3-17
ParticleFilter::ParticleFilter(int N, TopDownMapPolar* map, FilterParams& params) {
  std::random_device rd;
  gen_ = new std::mt19937(rd());
  num_gaussians_ = 1;
  map_ = map;
  params_ = params;
  max_num_particles_ = N;
  num_particles_ = 0;
  last_map_center_ = Eigen::Vector2i::Zero();
  if (map_->haveMap()) initializeParticles();
}

// the filter's parameters, the heading candidates (state_particle.cpp:197, :123-128) and the host set -> device
static bool upload_filter(ParticleFilter* f, const FilterParams& p, TopDownMapPolar* map, const std::vector<std::shared_ptr<StateParticle>>& particles) {
  if (!tdr() || particles.empty()) return false;
  tdr_filter_params fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.regularization = p.regularization; fp.force_on_map = p.force_on_map ? 1 : 0; fp.fixed_scale = p.fixed_scale;
  fp.scale_log_min = p.scale_log_min; fp.scale_log_max = p.scale_log_max; fp.num_classes = map->numClasses();
  for (int c = 0; c < fp.num_classes && c < 16; c++) fp.class_weights[c] = c < (int)p.class_weights.size() ? p.class_weights[c] : 1.f;
  if (!ok(tdr_pf_set_params(tdr(), &fp))) return false;
  std::vector<float> thetas;
  std::vector<int32_t> shifts;
  const int num_bins = 100;                                                  // rows of the polar scan images (top_down_render.cpp:115,530)
  for (float t = 0; t < 2 * M_PI; t += 2 * M_PI / 40) {
    int shift = static_cast<int>(std::round(t * num_bins / 2 / M_PI));
    while (shift >= num_bins) shift -= num_bins;
    while (shift < 0) shift += num_bins;
    thetas.push_back(t);
    shifts.push_back(shift);
  }
  if (!ok(tdr_pf_set_search(tdr(), thetas.data(), shifts.data(), (int)thetas.size()))) return false;
  std::vector<float> last_dist;
  std::vector<tdr_state> st = states_of(particles, last_dist);
  if (!ok(tdr_pf_set_states(tdr(), st.data(), last_dist.data(), (int64_t)st.size()))) return false;
  g_owner = f;
  return true;
}

// :19-84 — the same three constructions per particle from the shared engine; then the set goes to the device
void ParticleFilter::initializeParticles() {
  size_t num_at_scale = 1;
  if (params_.fixed_scale < 0) num_at_scale = 10; else scale_frozen_ = true;
  if (scale_frozen_ && params_.init_pos_m_x != std::numeric_limits<float>::infinity()) {
    Eigen::Vector2i map_center = map_->mapCenter();
    params_.init_pos_px_x = (params_.init_pos_m_x * params_.fixed_scale) + map_center.x();
    params_.init_pos_px_y = (params_.init_pos_m_y * params_.fixed_scale) + map_center.y();
    if (params_.init_pos_px_x < 0 || params_.init_pos_px_x >= map_->size()[0] || params_.init_pos_px_y < 0 || params_.init_pos_px_y >= map_->size()[1]) {
      ROS_WARN("[XView] No map received for input loc");
      return;
    }
    bool good_init = false;
    std::vector<int> here;
    for (int dx = -4; dx <= 4 && !good_init; dx++)
      for (int dy = -4; dy <= 4 && !good_init; dy++) {
        map_->getClassesAtPoint(Eigen::Vector2i(params_.init_pos_px_x + dx, params_.init_pos_px_y + dy), here);
        good_init = std::find(here.begin(), here.end(), 1) != here.end();
      }
    if (!good_init) { ROS_WARN("[XView] No road in map at init location"); return; }
  }
  for (int i = 0; i < max_num_particles_ / num_at_scale; i++) {
    StateParticle proto_part(gen_, map_, &params_);
    for (float scale = 0; scale < 1; scale += 1. / num_at_scale) {
      std::shared_ptr<StateParticle> particle = std::make_shared<StateParticle>(gen_, map_, &params_);
      if (params_.fixed_scale < 0) { particle->setState(proto_part.state()); particle->setScale(std::pow(10., scale)); }
      particles_.push_back(particle);
      new_particles_.push_back(std::make_shared<StateParticle>(gen_, map_, &params_));
    }
  }
  max_likelihood_particle_ = particles_[0];
  num_particles_ = particles_.size();
  weights_ = Eigen::Matrix<float, 1, Eigen::Dynamic>::Ones(num_particles_) / num_particles_;
  upload_filter(this, params_, map_, particles_);
  computeGMM();
  gmm_thread_ = new std::thread(std::bind(&ParticleFilter::gmmThread, this));
}

// the device set is this filter's, with the host mirror's current values (another filter may have used the context since)
static bool ensure_on_device(ParticleFilter* f, const FilterParams& p, TopDownMapPolar* map, const std::vector<std::shared_ptr<StateParticle>>& particles) {
  return g_owner == f || upload_filter(f, p, map, particles);
}
// the device set after a bulk call -> the staging area the per-particle members read from
static bool stage_device_set(size_t expect, const float* raw_weights) {
  int64_t n = 0;
  if (!ok(tdr_pf_count(tdr(), &n)) || n != (int64_t)expect) return false;
  g_stage_states.resize((size_t)n); g_stage_last_dist.resize((size_t)n);
  if (!ok(tdr_pf_get_states(tdr(), g_stage_states.data(), n)) || !ok(tdr_pf_get_last_dist(tdr(), g_stage_last_dist.data(), n))) return false;
  if (raw_weights) g_stage_weight.assign(raw_weights, raw_weights + n);
  g_cursor = 0;
  return true;
}

// :86-92 with StateParticle::propagate (:57-78): the reference's RNG calls in particle order, as STANDARD variates (the
// uniforms consumed do not depend on the standard deviation); the device applies `z * stddev + mean` and the motion
void ParticleFilter::propagate(Eigen::Vector2f& trans, float omega) {
  std::lock_guard<std::mutex> guard(particle_lock_);
  if (particles_.empty()) return;
  // host values first: ParticleFilter::updateMap / freezeScale / a harness may have edited the mirror
  upload_filter(this, params_, map_, particles_);
  std::vector<float> z(4 * particles_.size(), 0.f);
  for (size_t i = 0; i < particles_.size(); i++) {
    std::normal_distribution<float> disp_dist{0, 1}, theta_dist{0, 1};
    z[4 * i] = theta_dist(*gen_);
    z[4 * i + 1] = disp_dist(*gen_);
    z[4 * i + 2] = disp_dist(*gen_);
    if (!scale_frozen_) { std::normal_distribution<float> scale_dist{0, 1}; z[4 * i + 3] = scale_dist(*gen_); }
  }
  if (!ok(tdr_pf_propagate(tdr(), trans[0], trans[1], omega, scale_frozen_ ? 1 : 0, params_.pos_cov, params_.theta_cov, z.data(),
                           (int64_t)particles_.size()))) return;
  if (!stage_device_set(particles_.size(), nullptr)) return;
  for (auto& p : particles_) p->propagate(trans, omega, scale_frozen_);      // the reference's loop; each particle takes its result
}

// :94-189
void ParticleFilter::update(std::vector<Eigen::ArrayXXf>& top_down_scan, std::vector<Eigen::ArrayXXf>& /*top_down_geo*/, float res) {
  if (num_particles_ == 0) return;
  std::lock_guard<std::mutex> guard(particle_lock_);
  if (!upload_filter(this, params_, map_, particles_)) return;
  // the map (and its polar table) must be the one this filter scores against
  std::vector<Eigen::ArrayXXf> probe(map_->numClasses(), Eigen::ArrayXXf(top_down_scan[0].rows(), top_down_scan[0].cols()));
  Eigen::ArrayXXc probe_mask(top_down_scan[0].rows(), top_down_scan[0].cols());
  map_->getLocalMap(Eigen::Vector2f(0, 0), 1, res, probe, probe_mask);       // installs map + table on the device when they are not
  std::vector<float> scan;
  for (const auto& img : top_down_scan) scan.insert(scan.end(), img.data(), img.data() + img.size());
  if (!ok(tdr_scan_set_polar_images(tdr(), scan.data(), (int)top_down_scan[0].rows(), (int)top_down_scan[0].cols(), (int)top_down_scan.size()))) return;
  const size_t n = particles_.size();
  std::vector<float> raw(n);
  if (!ok(tdr_pf_score(tdr(), res, raw.data()))) return;                     // a9, a10 for the whole set
  if (!stage_device_set(n, raw.data())) return;
  std::vector<Eigen::ArrayXXf> no_geo;
  for (auto& p : particles_) p->computeWeight(top_down_scan, no_geo, res);   // the reference's loop (:104-105); each takes its weight
  int64_t arg = 0;
  if (!ok(tdr_pf_normalize(tdr(), &arg, nullptr))) return;                   // a11
  weights_ = Eigen::VectorXf(n);
  if (!ok(tdr_pf_get_weights(tdr(), weights_.data(), (int64_t)n))) return;
  max_likelihood_particle_ = particles_[arg];
  int last_num_particles = num_particles_;                                   // :151-158
  num_particles_ = 0;
  for (const auto& cov : covs_) {
    Eigen::Vector2cf eig = cov.block<2, 2>(0, 0).eigenvalues();
    num_particles_ += static_cast<int>(sqrt(eig[0].real()) * sqrt(eig[1].real()));
  }
  num_particles_ = std::min(std::max(num_particles_, 3 * last_num_particles / 4 + 10), max_num_particles_);
  if ((size_t)num_particles_ < new_particles_.size()) new_particles_.resize(num_particles_);
  while (new_particles_.size() < (size_t)num_particles_) new_particles_.push_back(std::make_shared<StateParticle>(gen_, map_, &params_, false));
  std::uniform_real_distribution<float> shift_dist(0., 1.);
  float shift = shift_dist(*gen_);                                           // the ONE draw of :172-173
  std::vector<int32_t> idx((size_t)num_particles_);
  if (!ok(tdr_pf_resample(tdr(), shift, num_particles_, idx.data()))) return;   // a12: indices from the device
  for (int i = 0; i < num_particles_; i++) new_particles_[i]->setState(particles_[idx[i]]->state());   // :185, the mirror follows
  particles_.swap(new_particles_);
}

// :191-236 — the sums run on the device; the ML particle is a host object, as in the reference
void ParticleFilter::meanLikelihood(Eigen::Vector4f& mean_state) {
  mean_state = Eigen::Vector4f::Zero();
  if (particles_.empty() || !ensure_on_device(this, params_, map_, particles_)) return;
  ok(tdr_pf_pose(tdr(), mean_state.data(), nullptr, nullptr, nullptr));
}
void ParticleFilter::computeMeanCov(Eigen::Matrix4f& cov) {
  cov.setZero();
  if (num_particles_ < 1 || !ensure_on_device(this, params_, map_, particles_)) return;
  float mean[4];
  ok(tdr_pf_pose(tdr(), mean, cov.data(), nullptr, nullptr));
}
void ParticleFilter::maxLikelihood(Eigen::Vector4f& state) { state = max_likelihood_particle_->mlState(); }
void ParticleFilter::computeCov(Eigen::Matrix4f& cov) {
  cov.setZero();
  Eigen::Vector4f ml = max_likelihood_particle_->mlState();
  for (const auto& particle : particles_) {                                  // N x 16 flops on the host mirror; off the hot path
    Eigen::Vector4f state = particle->mlState() - ml;
    while (state[2] > M_PI) state[2] -= 2 * M_PI;
    while (state[2] < -M_PI) state[2] += 2 * M_PI;
    cov += state * state.transpose();
  }
  cov /= particles_.size() - 1;
}

void ParticleFilter::getGMM(std::vector<Eigen::Vector3f>& means, std::vector<Eigen::Matrix3f>& covs) {
  std::lock_guard<std::mutex> guard(gmm_lock_);
  means = means_;
  covs = covs_;
}
void ParticleFilter::gmmThread() {
  while (true) {
    computeGMM();
    std::this_thread::sleep_for(std::chrono::milliseconds(1000));
  }
}
// :252-318 — the EM input comes off the host mirror (the same matrix tdr_pf_gmm_samples builds, without a device call from
// the GMM thread); the fit is OpenCV's cv::ml::EM, tried with one cluster more and one fewer as the reference does
void ParticleFilter::computeGMM() {
  cv::Mat samples;
  {
    std::lock_guard<std::mutex> guard(particle_lock_);
    const size_t n = particles_.size();
    num_gaussians_ = std::min(static_cast<int>(n / 20) + 1, num_gaussians_);
    const int num_samples = std::min(1000, static_cast<int>(n));
    samples = cv::Mat(num_samples, 4, CV_64F);
    for (int i = 0; i < num_samples; i++) {
      Eigen::Vector3f s = particles_[std::min<int>(n - 1, i * n / num_samples)]->mlState().head<3>();
      double* row = &samples.at<double>(i, 0);
      row[0] = s[0]; row[1] = s[1]; row[2] = 50 * cos(s[2]); row[3] = 50 * sin(s[2]);
    }
  }
  cv::Ptr<cv::ml::EM> em = cv::ml::EM::create();
  em->setCovarianceMatrixType(cv::ml::EM::COV_MAT_GENERIC);
  cv::Mat likelihoods, labels;
  auto mean_likelihood = [&](int clusters) -> float {
    em->setClustersNumber(clusters);
    em->trainEM(samples, likelihoods, labels);
    return cv::mean(likelihoods)[0];
  };
  const float base = mean_likelihood(num_gaussians_);
  int dir = 0;
  if (num_gaussians_ * 50 < num_particles_ && base + 0.3 < mean_likelihood(num_gaussians_ + 1)) dir = 1;
  if (num_gaussians_ > 1 && base - 0.3 < mean_likelihood(num_gaussians_ - 1)) dir = -1;
  num_gaussians_ += dir;
  mean_likelihood(num_gaussians_);
  cv::Mat mu = em->getMeans();
  std::vector<cv::Mat> sigma;
  em->getCovs(sigma);
  std::lock_guard<std::mutex> guard(gmm_lock_);
  means_.clear();
  covs_.clear();
  for (size_t k = 0; k < sigma.size(); k++) {
    means_.push_back(Eigen::Vector3f(mu.at<double>(k, 0), mu.at<double>(k, 1), atan2(mu.at<double>(k, 3), mu.at<double>(k, 2))));
    Eigen::Matrix3f cov;
    cov << sigma[k].at<double>(0, 0), sigma[k].at<double>(0, 1), 0, sigma[k].at<double>(1, 0), sigma[k].at<double>(1, 1), 0, 0, 0, 1;
    covs_.push_back(cov);
  }
}

// :320-341
void ParticleFilter::updateMap(const cv::Mat& map, const Eigen::Vector2i& map_center) {
  map_->updateMap(map, map_center);
  Eigen::Vector2i delta = map_center - last_map_center_;
  {
    std::lock_guard<std::mutex> guard(particle_lock_);
    for (auto& particle : particles_) {
      State s = particle->state();
      s.init_x_px += delta[0];
      s.init_y_px += delta[1];
      particle->setState(s);
      particle->updateSize();
    }
    g_owner = nullptr;                                                       // the mirror is ahead of the device
  }
  last_map_center_ = map_center;
  if (num_particles_ == 0) initializeParticles();
}
// :343-357
void ParticleFilter::freezeScale() {
  if (scale_frozen_) return;
  float geo_mean = 1;
  for (const auto& p : particles_) geo_mean *= std::pow(p->state().scale, 1. / particles_.size());
  for (auto& p : particles_) p->setScale(geo_mean);
  scale_frozen_ = true;
  g_owner = nullptr;
}
// :358-366
float ParticleFilter::scale() const {
  if (params_.fixed_scale > 0) return params_.fixed_scale;
  if (scale_frozen_) return particles_[0]->state().scale;
  return -1;
}
int ParticleFilter::numParticles() const { return num_particles_; }
