// adapters/particle_filter_adapter.cpp — the bodies that REPLACE src/particle_filter.cpp and src/state_particle.cpp of the reference.
// The class declarations are the reference's own, unchanged (include/top_down_render/particle_filter.h:22-73,
// state_particle.h:9-71).  Scoring, normalisation, resampling, propagation and the pose run on the device through the C ABI
// (include/tdr.h); what stays host code is what is sequential on the ONE shared std::mt19937 in the reference too — the
// rejection sampling of the StateParticle constructor, the standard-normal variates of propagate, the single uniform of
// update — so a run with the same seed consumes the engine exactly as the reference does.
//
// particles_ / new_particles_ / weights_ — the members the header declares — are a host MIRROR of the device set, refreshed
// LAZILY: the node's per-scan calls, propagate and update, move no particle data over PCIe (16 B of
// standard variates per particle go up for propagate, or nothing with TDR_ADAPTER_DEVICE_RNG=1); the pose members read
// the device; visualize, the map shift, freezeScale and a test harness pull the states when they read them
// (tdr_adapter_sync_mirror).  The headers give ParticleFilter no access to StateParticle's private fields (weight_,
// last_dist_), so the refresh goes through StateParticle's own members: their bodies here take one particle's share of a
// staged result, in particle order.  What the adapter needs beyond the declared members (is the mirror behind the device?
// which arg-max?) lives in a side table keyed by the filter — the class declaration stays the reference's.
#include <cstring>
#include <map>
#include <mutex>
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
// per thread: the spinner thread and the GMM thread never share a staging area (and both hold particle_lock_ anyway)
struct Stage {
  std::vector<tdr_state> states;
  std::vector<float> last_dist, weight;
  size_t cursor = 0;
};
thread_local Stage g_stage;
}  // namespace
void StateParticle::propagate(Eigen::Vector2f&, float, bool) {
  const tdr_state& t = g_stage.states[g_stage.cursor];
  state_.init_x_px = t.init_x_px; state_.init_y_px = t.init_y_px;
  state_.dx_m = t.dx_m; state_.dy_m = t.dy_m; state_.theta = t.theta; state_.scale = t.scale; state_.have_init = t.have_init != 0;
  last_dist_ = g_stage.last_dist[g_stage.cursor++];
}
void StateParticle::computeWeight(std::vector<Eigen::ArrayXXf>&, std::vector<Eigen::ArrayXXf>&, float) {
  const tdr_state& t = g_stage.states[g_stage.cursor];
  state_.theta = t.theta;                                                     // the heading search's choice (:195-206)
  state_.have_init = t.have_init != 0;
  weight_ = g_stage.weight[g_stage.cursor++];
}

// ============================ ParticleFilter (particle_filter.cpp) ==========================================================
namespace {
// what the adapter knows about a filter beyond its declared members
struct Side {
  enum Where { IN_SYNC, HOST_AHEAD, DEVICE_AFTER_PROPAGATE, DEVICE_AFTER_UPDATE } where = HOST_AHEAD;
  int64_t argmax = 0;                 // arg-max of the last normalisation (max_likelihood_particle_ = scored set [argmax])
  bool have_params = false;
  tdr_filter_params params{};         // what the device holds
  size_t scored = 0;                  // size of the set the last update scored
  // the four pose members of one scan (top_down_render.cpp:423-431 calls them back to back) share ONE device pass
  bool have_pose = false;
  float mean[4], cov_mean[16], ml[4], cov_ml[16];
  bool have_seed = false;             // TDR_ADAPTER_DEVICE_RNG: Philox key (from the shared engine, once) and step counter
  uint64_t rng_seed = 0, steps = 0;
};
// heap objects that are never destroyed: the detached GMM thread may still run while the process exits
std::mutex& g_side_lock = *new std::mutex;
std::map<const ParticleFilter*, Side>& g_side = *new std::map<const ParticleFilter*, Side>;
// the device holds ONE particle set: the filter that used it last
ParticleFilter* g_owner = nullptr;
Side& side_of(const ParticleFilter* f) { std::lock_guard<std::mutex> g(g_side_lock); return g_side[f]; }
bool device_rng() { static const bool on = [] { const char* e = getenv("TDR_ADAPTER_DEVICE_RNG"); return e && atoi(e) != 0; }(); return on; }

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

// the filter's parameters and the heading candidates (state_particle.cpp:197, :123-128) -> device, when they changed
static bool upload_params(Side& sd, const FilterParams& p, TopDownMapPolar* map) {
  tdr_filter_params fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.regularization = p.regularization; fp.force_on_map = p.force_on_map ? 1 : 0; fp.fixed_scale = p.fixed_scale;
  fp.scale_log_min = p.scale_log_min; fp.scale_log_max = p.scale_log_max; fp.num_classes = map->numClasses();
  for (int c = 0; c < fp.num_classes && c < 16; c++) fp.class_weights[c] = c < (int)p.class_weights.size() ? p.class_weights[c] : 1.f;
  if (sd.have_params && std::memcmp(&fp, &sd.params, sizeof(fp)) == 0) return true;      // the fp16 / u8 map copies stay valid
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
  if (!ok(tdr_pf_keep_raw_weights(tdr(), 1))) return false;                  // StateParticle::weight() of a refreshed mirror
  sd.params = fp; sd.have_params = true;
  return true;
}

// device set -> host mirror of ONE filter (its private members, handed in by a member function), if the device is ahead.
// After an update the reference's vectors read: particles_ = the resampled set, new_particles_ = the set that was scored
// (raw weight, searched heading), weights_ = the normalised weights, max_likelihood_particle_ = scored set [arg-max].
static bool pull_mirror(Side& sd, std::mt19937* gen, TopDownMapPolar* map, FilterParams* params,
                        std::vector<std::shared_ptr<StateParticle>>& particles, std::vector<std::shared_ptr<StateParticle>>& new_particles,
                        Eigen::VectorXf& weights, std::shared_ptr<StateParticle>& ml_particle) {
  if (sd.where != Side::DEVICE_AFTER_PROPAGATE && sd.where != Side::DEVICE_AFTER_UPDATE) return true;
  if (!tdr()) return false;
  int64_t n = 0;
  if (!ok(tdr_pf_count(tdr(), &n)) || n <= 0) return false;
  Eigen::Vector2f no_trans(0, 0);
  std::vector<Eigen::ArrayXXf> none;
  Stage& sg = g_stage;
  if (sd.where == Side::DEVICE_AFTER_UPDATE) {
    // the scored set first: states as the search left them, last_dist, raw weights -> new_particles_
    const int64_t m = (int64_t)sd.scored;
    sg.states.resize((size_t)m); sg.last_dist.resize((size_t)m); sg.weight.resize((size_t)m);
    if (!ok(tdr_pf_get_prev_states(tdr(), sg.states.data(), sg.last_dist.data(), m)) || !ok(tdr_pf_get_raw_weights(tdr(), sg.weight.data(), m))) return false;
    if (new_particles.size() > (size_t)m) new_particles.resize((size_t)m);
    while (new_particles.size() < (size_t)m) new_particles.push_back(std::make_shared<StateParticle>(gen, map, params, false));
    sg.cursor = 0;
    for (auto& p : new_particles) p->propagate(no_trans, 0.f, true);          // every State field + last_dist_
    sg.cursor = 0;
    for (auto& p : new_particles) p->computeWeight(none, none, 0.f);          // weight_ (and heading / have_init again)
    weights = Eigen::VectorXf(m);
    if (!ok(tdr_pf_get_weights(tdr(), weights.data(), m))) return false;
    ml_particle = new_particles[(size_t)sd.argmax];
  }
  sg.states.resize((size_t)n); sg.last_dist.resize((size_t)n);
  if (!ok(tdr_pf_get_states(tdr(), sg.states.data(), n)) || !ok(tdr_pf_get_last_dist(tdr(), sg.last_dist.data(), n))) return false;
  if (particles.size() > (size_t)n) particles.resize((size_t)n);
  while (particles.size() < (size_t)n) particles.push_back(std::make_shared<StateParticle>(gen, map, params, false));
  sg.cursor = 0;
  for (auto& p : particles) p->propagate(no_trans, 0.f, true);
  sd.where = Side::IN_SYNC;
  return true;
}

// host mirror -> device, when the mirror is ahead (initialisation, map shift, freezeScale, a harness edit) or the context
// last served another filter
static bool push_mirror(ParticleFilter* f, Side& sd, const FilterParams& p, TopDownMapPolar* map,
                        const std::vector<std::shared_ptr<StateParticle>>& particles) {
  if (!tdr() || particles.empty()) return false;
  if (!upload_params(sd, p, map)) return false;
  if (g_owner == f && sd.where != Side::HOST_AHEAD) return true;
  std::vector<float> last_dist;
  std::vector<tdr_state> st = states_of(particles, last_dist);
  if (!ok(tdr_pf_set_states(tdr(), st.data(), last_dist.data(), (int64_t)st.size()))) return false;
  g_owner = f;
  sd.where = Side::IN_SYNC; sd.have_pose = false;
  return true;
}

// Before a filter touches the device: if the context last served ANOTHER filter whose mirror is behind the device, that
// filter's mirror is refreshed first.  A macro because only code inside a ParticleFilter member may reach another
// ParticleFilter's private members, and the class declaration is not ours to extend.
#define take_device()                                                                                                   \
  do {                                                                                                                  \
    if (g_owner && g_owner != this) {                                                                                   \
      ParticleFilter* o_ = g_owner;                                                                                     \
      Side& so_ = side_of(o_);                                                                                          \
      pull_mirror(so_, o_->gen_, o_->map_, &o_->params_, o_->particles_, o_->new_particles_, o_->weights_, o_->max_likelihood_particle_); \
      so_.where = Side::HOST_AHEAD;                                                                                     \
      g_owner = nullptr;                                                                                                \
    }                                                                                                                   \
  } while (0)

// for code that edits a filter's particles_ behind the class's back (the test harness does, through its private-access
// build): the next propagate / update uploads the mirror again
extern "C" void tdr_adapter_mark_host_ahead(const void* filter) {
  Side& sd = side_of(static_cast<const ParticleFilter*>(filter));
  sd.where = Side::HOST_AHEAD; sd.have_pose = false;
}

// :19-84 — the same three constructions per particle from the shared engine; then the set goes to the device
void ParticleFilter::initializeParticles() {
  tdr_adapter::Guard dev_guard;
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
  { Side& sd = side_of(this); sd.where = Side::HOST_AHEAD; take_device(); push_mirror(this, sd, params_, map_, particles_); }
  computeGMM();
  gmm_thread_ = new std::thread(std::bind(&ParticleFilter::gmmThread, this));
}

// :86-92 with StateParticle::propagate (:57-78).  The reference's RNG calls in particle order, as STANDARD variates (the
// uniforms consumed do not depend on the standard deviation): the device applies `z * stddev + mean` and the motion, and
// the shared engine ends where the reference's would.  TDR_ADAPTER_DEVICE_RNG=1 draws on the device instead (no host
// loop over the particles at all; the engine is then not advanced by propagate).  No state comes back.
void ParticleFilter::propagate(Eigen::Vector2f& trans, float omega) {
  tdr_adapter::Guard dev_guard;
  std::lock_guard<std::mutex> guard(particle_lock_);
  if (particles_.empty()) return;
  Side& sd = side_of(this);
  take_device();
  if (!push_mirror(this, sd, params_, map_, particles_)) return;
  const size_t n = particles_.size();
  if (device_rng()) {
    if (!sd.have_seed) { sd.rng_seed = ((uint64_t)(*gen_)() << 32) | (uint64_t)(*gen_)(); sd.have_seed = true; }   // two engine outputs, once
    if (!ok(tdr_pf_propagate_rng(tdr(), trans[0], trans[1], omega, scale_frozen_ ? 1 : 0, params_.pos_cov, params_.theta_cov,
                                 sd.rng_seed, ++sd.steps, nullptr))) return;
  } else {
    std::vector<float> z(4 * n, 0.f);
    for (size_t i = 0; i < n; i++) {
      std::normal_distribution<float> disp_dist{0, 1}, theta_dist{0, 1};
      z[4 * i] = theta_dist(*gen_);
      z[4 * i + 1] = disp_dist(*gen_);
      z[4 * i + 2] = disp_dist(*gen_);
      if (!scale_frozen_) { std::normal_distribution<float> scale_dist{0, 1}; z[4 * i + 3] = scale_dist(*gen_); }
    }
    if (!ok(tdr_pf_propagate(tdr(), trans[0], trans[1], omega, scale_frozen_ ? 1 : 0, params_.pos_cov, params_.theta_cov, z.data(), (int64_t)n))) return;
  }
  sd.where = Side::DEVICE_AFTER_PROPAGATE; sd.have_pose = false;
}

// :94-189 on the resident set: score, normalise, the adaptive particle count, ONE uniform from the shared engine, resample
void ParticleFilter::update(std::vector<Eigen::ArrayXXf>& top_down_scan, std::vector<Eigen::ArrayXXf>& /*top_down_geo*/, float res) {
  tdr_adapter::Guard dev_guard;
  if (num_particles_ == 0) return;
  std::lock_guard<std::mutex> guard(particle_lock_);
  Side& sd = side_of(this);
  take_device();
  if (!push_mirror(this, sd, params_, map_, particles_)) return;
  // the map (and its polar table) must be the one this filter scores against
  std::vector<Eigen::ArrayXXf> probe(map_->numClasses(), Eigen::ArrayXXf(top_down_scan[0].rows(), top_down_scan[0].cols()));
  Eigen::ArrayXXc probe_mask(top_down_scan[0].rows(), top_down_scan[0].cols());
  map_->getLocalMap(Eigen::Vector2f(0, 0), 1, res, probe, probe_mask);       // installs map + table on the device when they are not
  std::vector<float> scan;
  for (const auto& img : top_down_scan) scan.insert(scan.end(), img.data(), img.data() + img.size());
  if (!ok(tdr_scan_set_polar_images(tdr(), scan.data(), (int)top_down_scan[0].rows(), (int)top_down_scan[0].cols(), (int)top_down_scan.size()))) return;
  int64_t n = 0;
  if (!ok(tdr_pf_count(tdr(), &n))) return;
  if (!ok(tdr_pf_score(tdr(), res, nullptr))) return;                        // a9, a10 for the whole set; the weights stay on the device
  int64_t arg = 0;
  if (!ok(tdr_pf_normalize(tdr(), &arg, nullptr))) return;                   // a11
  int last_num_particles = num_particles_;                                   // :151-158
  num_particles_ = 0;
  for (const auto& cov : covs_) {
    Eigen::Vector2cf eig = cov.block<2, 2>(0, 0).eigenvalues();
    num_particles_ += static_cast<int>(sqrt(eig[0].real()) * sqrt(eig[1].real()));
  }
  num_particles_ = std::min(std::max(num_particles_, 3 * last_num_particles / 4 + 10), max_num_particles_);
  std::uniform_real_distribution<float> shift_dist(0., 1.);
  float shift = shift_dist(*gen_);                                           // the ONE draw of :172-173
  if (!ok(tdr_pf_resample(tdr(), shift, num_particles_, nullptr))) return;   // a12 + the state gather (:185) on the device
  sd.argmax = arg; sd.scored = (size_t)n;
  sd.where = Side::DEVICE_AFTER_UPDATE; sd.have_pose = false;
}

// :191-236 — the sums run on the device, the max-likelihood state too (the arg-max of the last normalisation); the four
// members of one scan share one device pass (Side::have_pose)
static bool device_pose(Side& sd, bool with_ml) {
  if (sd.have_pose) return true;
  if (!ok(tdr_pf_pose(tdr(), sd.mean, sd.cov_mean, with_ml ? sd.ml : nullptr, with_ml ? sd.cov_ml : nullptr))) return false;
  sd.have_pose = with_ml;
  return true;
}
void ParticleFilter::meanLikelihood(Eigen::Vector4f& mean_state) {
  tdr_adapter::Guard dev_guard;
  mean_state = Eigen::Vector4f::Zero();
  std::lock_guard<std::mutex> guard(particle_lock_);
  Side& sd = side_of(this);
  take_device();
  if (particles_.empty() || !push_mirror(this, sd, params_, map_, particles_)) return;
  if (device_pose(sd, sd.where == Side::DEVICE_AFTER_UPDATE)) std::memcpy(mean_state.data(), sd.mean, sizeof(sd.mean));
}
void ParticleFilter::computeMeanCov(Eigen::Matrix4f& cov) {
  tdr_adapter::Guard dev_guard;
  cov.setZero();
  std::lock_guard<std::mutex> guard(particle_lock_);
  Side& sd = side_of(this);
  take_device();
  if (num_particles_ < 1 || !push_mirror(this, sd, params_, map_, particles_)) return;
  if (device_pose(sd, sd.where == Side::DEVICE_AFTER_UPDATE)) std::memcpy(cov.data(), sd.cov_mean, sizeof(sd.cov_mean));
}
void ParticleFilter::maxLikelihood(Eigen::Vector4f& state) {
  tdr_adapter::Guard dev_guard;
  std::lock_guard<std::mutex> guard(particle_lock_);
  Side& sd = side_of(this);
  if (sd.where == Side::DEVICE_AFTER_UPDATE && g_owner == this) { if (device_pose(sd, true)) std::memcpy(state.data(), sd.ml, sizeof(sd.ml)); return; }
  state = max_likelihood_particle_->mlState();
}
void ParticleFilter::computeCov(Eigen::Matrix4f& cov) {
  tdr_adapter::Guard dev_guard;
  cov.setZero();
  std::lock_guard<std::mutex> guard(particle_lock_);
  Side& sd = side_of(this);
  if (sd.where == Side::DEVICE_AFTER_UPDATE && g_owner == this) { if (device_pose(sd, true)) std::memcpy(cov.data(), sd.cov_ml, sizeof(sd.cov_ml)); return; }
  Eigen::Vector4f ml = max_likelihood_particle_->mlState();
  for (const auto& particle : particles_) {                                  // the mirror is current: the reference's loop
    Eigen::Vector4f state = particle->mlState() - ml;
    while (state[2] > M_PI) state[2] -= 2 * M_PI;
    while (state[2] < -M_PI) state[2] += 2 * M_PI;
    cov += state * state.transpose();
  }
  cov /= particles_.size() - 1;
}

// :367-421 — the drawing walks every particle of the host mirror: this is where a node that visualises pays the 28 B per
// particle; a harness calls it (with any image) to read a current mirror.  The drawing itself needs OpenCV's imgproc.
void ParticleFilter::visualize(cv::Mat& img) {
  tdr_adapter::Guard dev_guard;
  {
    std::lock_guard<std::mutex> guard(particle_lock_);
    Side& sd = side_of(this);
    if (g_owner == this) pull_mirror(sd, gen_, map_, &params_, particles_, new_particles_, weights_, max_likelihood_particle_);
  }
#ifdef TDR_ADAPTER_DRAW
  tdr_adapter_draw(img, particles_, means_, covs_, gmm_lock_, max_likelihood_particle_);   // the reference's drawing code, unchanged
#else
  (void)img;
#endif
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
    tdr_adapter::Guard dev_guard;
    std::lock_guard<std::mutex> guard(particle_lock_);
    // once a second the GMM thread reads a strided sample of the set: the mirror is refreshed here if the device is ahead
    if (g_owner == this) pull_mirror(side_of(this), gen_, map_, &params_, particles_, new_particles_, weights_, max_likelihood_particle_);
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
    tdr_adapter::Guard dev_guard;
    std::lock_guard<std::mutex> guard(particle_lock_);
    Side& sd = side_of(this);
    if (g_owner == this) pull_mirror(sd, gen_, map_, &params_, particles_, new_particles_, weights_, max_likelihood_particle_);
    for (auto& particle : particles_) {
      State s = particle->state();
      s.init_x_px += delta[0];
      s.init_y_px += delta[1];
      particle->setState(s);
      particle->updateSize();
    }
    sd.where = Side::HOST_AHEAD; sd.have_pose = false;                       // the mirror is ahead of the device
  }
  last_map_center_ = map_center;
  if (num_particles_ == 0) initializeParticles();
}
// :343-357
void ParticleFilter::freezeScale() {
  if (scale_frozen_) return;
  tdr_adapter::Guard dev_guard;
  std::lock_guard<std::mutex> guard(particle_lock_);
  Side& sd = side_of(this);
  if (g_owner == this) pull_mirror(sd, gen_, map_, &params_, particles_, new_particles_, weights_, max_likelihood_particle_);
  float geo_mean = 1;
  for (const auto& p : particles_) geo_mean *= std::pow(p->state().scale, 1. / particles_.size());
  for (auto& p : particles_) p->setScale(geo_mean);
  scale_frozen_ = true;
  sd.where = Side::HOST_AHEAD; sd.have_pose = false;
}
// :358-366
float ParticleFilter::scale() const {
  if (params_.fixed_scale > 0) return params_.fixed_scale;
  if (scale_frozen_) return particles_[0]->state().scale;                    // frozen: every particle holds it, the device changes it no more
  return -1;
}
int ParticleFilter::numParticles() const { return num_particles_; }
