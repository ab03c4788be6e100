// host_demo.cpp — drives the C++ host mirror (top_down_renderer_b200/host/tdr_host.hpp) through one scan step the
// way TopDownRender::takeStep does (top_down_render.cpp:505-572): map update, polar rasterise, particle init,
// propagate, update, pose.  Inputs / outputs are raw little-endian files in argv[1] so that tests/test_host_cpp.py
// can check them against the oracle.
#include <fstream>
#include <iostream>
#include <limits>

#include "../../top_down_renderer_b200/host/tdr_host.hpp"

using namespace tdrhost;

template <typename T> static std::vector<T> rd(const std::string& p) {
  std::ifstream f(p, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "cannot read " << p << "\n"; exit(2); }
  size_t n = f.tellg(); f.seekg(0);
  std::vector<T> v(n / sizeof(T));
  f.read(reinterpret_cast<char*>(v.data()), n);
  return v;
}
template <typename T> static void wr(const std::string& p, const T* d, size_t n) {
  std::ofstream f(p, std::ios::binary);
  f.write(reinterpret_cast<const char*>(d), n * sizeof(T));
}

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: host_demo <dir>\n"; return 2; }
  const std::string dir = std::string(argv[1]) + "/";
  auto meta = rd<int32_t>(dir + "meta.i32");       // H, W, C, N, seed, n_theta, n_r
  auto fmeta = rd<float>(dir + "meta.f32");        // res, init_x, init_y, init_cov, init_theta_deg, init_theta_cov
  const int H = meta[0], W = meta[1], C = meta[2], N = meta[3], seed = meta[4], n_theta = meta[5], n_r = meta[6];
  auto img = rd<uint8_t>(dir + "class_image.u8");
  auto pts = rd<PointXYZI>(dir + "points.f32");
  if (!ctx()) return 3;                             // no GPU: nothing to demonstrate (and no CPU fallback)

  std::vector<int> lut(256, -1);
  for (int c = 0; c < C; c++) lut[c] = c;
  TopDownMapPolar::Params mp; mp.flatten_lut = lut; mp.num_classes = C; mp.resolution = 1;
  TopDownMapPolar map(mp);
  map.samplePtsPolar(n_theta, n_r, (float)(2 * M_PI / n_theta));
  map.updateMap(img.data(), H, W, W, Vector2i{W / 2, H / 2});
  if (!map.haveMap()) return 4;
  // the map cache (top_down_map.cpp:226-286): written in the reference's .eig format, then loaded back — from here on
  // the filter runs on the CACHED distance fields (tdr_map_set_dist_layers), which must change nothing
  map.saveCachedMaps(dir, "demo_map");
  if (!map.loadCacheMetaData(dir, "demo_map") || map.loadCacheMetaData(dir, "another_map")) return 5;
  if (!map.loadCachedMaps(dir) || !map.haveMap()) return 6;

  ScanRendererPolar renderer(lut);
  std::vector<ArrayXXf> top_down(C, ArrayXXf(n_theta, n_r)), geo;
  renderer.renderSemanticTopDown(pts, fmeta[0], (float)(2 * M_PI / n_theta), top_down);
  for (int c = 0; c < C; c++) wr(dir + "scan_" + std::to_string(c) + ".f32", top_down[c].data(), (size_t)n_theta * n_r);

  // the Cartesian twins through the base classes (non-virtual name hiding, as in the reference): the refine_map-style
  // raster of the scan and the node's debug view of the map around the start pose (top_down_render.cpp:315)
  {
    ScanRenderer& cart = renderer;
    std::vector<ArrayXXf> cart_imgs(C, ArrayXXf(48, 64));
    cart.renderSemanticTopDown(pts, 1.5f, cart_imgs);
    for (int c = 0; c < C; c++) wr(dir + "cart_scan_" + std::to_string(c) + ".f32", cart_imgs[c].data(), (size_t)48 * 64);
    TopDownMap* base = &map;
    std::vector<ArrayXXf> cart_local(C, ArrayXXf(30, 40));
    ArrayXXc cart_mask(30, 40);
    Vector2f cc; cc.x = fmeta[1]; cc.y = fmeta[2];
    base->getLocalMap(cc, 0.3f, 2.5f, cart_local, cart_mask);
    for (int c = 0; c < C; c++) wr(dir + "cart_local_" + std::to_string(c) + ".f32", cart_local[c].data(), (size_t)30 * 40);
    wr(dir + "cart_mask.u8", cart_mask.data(), (size_t)30 * 40);
  }

  FilterParams fp; fp.regularization = 0.7f; fp.pos_cov = 0.15f; fp.theta_cov = 0.004f; fp.fixed_scale = 2.0f;
  fp.init_pos_px_x = fmeta[1]; fp.init_pos_px_y = fmeta[2]; fp.init_pos_px_cov = fmeta[3];
  fp.init_pos_deg_theta = fmeta[4]; fp.init_pos_deg_cov = fmeta[5];
  fp.init_pos_m_x = fp.init_pos_m_y = std::numeric_limits<float>::infinity();   // "unset", as the node passes it (top_down_render.cpp:214-222)
  fp.class_weights.assign(C, 1.f);
  ParticleFilter filter(N, &map, fp, (uint32_t)seed);
  Vector2f trans; trans.x = 0.4f; trans.y = 0.05f;
  filter.propagate(trans, 0.01f);
  wr(dir + "states_before.bin", filter.states().data(), filter.states().size());
  wr(dir + "last_dist.f32", filter.lastDist().data(), filter.lastDist().size());
  filter.update(top_down, geo, fmeta[0]);
  const float u = filter.lastUniform();
  wr(dir + "u.f32", &u, 1);
  auto w = filter.weights();
  wr(dir + "weights_norm.f32", w.data(), w.size());
  wr(dir + "states_after.bin", filter.states().data(), filter.states().size());
  float mean[4], cov[16], ml[4];
  filter.meanLikelihood(mean); filter.computeMeanCov(cov); filter.maxLikelihood(ml);
  wr(dir + "mean.f32", mean, 4); wr(dir + "cov.f32", cov, 16); wr(dir + "ml.f32", ml, 4);
  // ActiveLocalizer (dead code in the reference, kept linkable): most discriminative relative position for three guesses
  ActiveLocalizer al(&map);
  std::vector<Vector3f> preds(3);
  preds[0].x = W * 0.3f; preds[0].y = H * 0.4f; preds[0].z = 0.2f;
  preds[1].x = W * 0.6f; preds[1].y = H * 0.5f; preds[1].z = -1.1f;
  preds[2].x = W * 0.5f; preds[2].y = H * 0.7f; preds[2].z = 2.4f;
  const Vector2f rel = al.getBestRelPos(preds);
  const float relv[2] = {rel.x, rel.y};
  wr(dir + "active_rel.f32", relv, 2);
  std::vector<ArrayXXf> geo_local(2, ArrayXXf(n_theta, n_r));
  Vector2f gc; gc.x = W * 0.5f; gc.y = H * 0.5f;
  map.getLocalGeoMap(gc, 2.0f, geo_local);
  wr(dir + "geo_local0.f32", geo_local[0].data(), (size_t)n_theta * n_r);
  // ---- the rest of the ParticleFilter interface (particle_filter.h:22-41); the first filter is not touched again ----
  const float sc_fixed = filter.scale();                                       // fixed scale: the parameter
  // (a) ParticleFilter::updateMap: a new aerial map whose centre moved by (+3, -2) px shifts every init position
  Vector2i c2; c2.x = W / 2 + 3; c2.y = H / 2 - 2;
  ParticleFilter shifted(64, &map, fp, (uint32_t)seed + 1);
  shifted.updateMap(img.data(), H, W, W, Vector2i{W / 2, H / 2});               // first map message: delta from (0, 0)
  wr(dir + "shift_before.bin", shifted.states().data(), shifted.states().size());
  shifted.updateMap(img.data(), H, W, W, c2);
  wr(dir + "shift_after.bin", shifted.states().data(), shifted.states().size());
  // (b) free scale: ten scales per prototype, scale() = -1 until freezeScale locks the geometric mean
  FilterParams ff = fp; ff.fixed_scale = -1; ff.init_pos_px_x = ff.init_pos_px_y = -1;
  ff.init_pos_deg_theta = std::numeric_limits<float>::infinity();
  ParticleFilter free_scale(120, &map, ff, (uint32_t)seed + 2);
  const float sc_free = free_scale.scale();
  wr(dir + "free_before.bin", free_scale.states().data(), free_scale.states().size());
  Vector2f t2; t2.x = 0.3f; t2.y = -0.1f;
  free_scale.propagate(t2, -0.02f);                                            // scale jitter: four variates per particle
  wr(dir + "free_propagated.bin", free_scale.states().data(), free_scale.states().size());
  free_scale.freezeScale();
  wr(dir + "free_frozen.bin", free_scale.states().data(), free_scale.states().size());
  const float sc_frozen = free_scale.scale();
  // (c) a metric initial position relative to the map centre (particle_filter.cpp:27-54), on the road and off the map
  FilterParams fm = fp; fm.init_pos_px_x = fm.init_pos_px_y = -1;
  fm.init_pos_m_x = (fmeta[1] - c2.x) / fp.fixed_scale; fm.init_pos_m_y = (fmeta[2] - c2.y) / fp.fixed_scale;
  ParticleFilter metric(48, &map, fm, (uint32_t)seed + 3);
  wr(dir + "metric_states.bin", metric.states().data(), metric.states().size());
  FilterParams fo = fm; fo.init_pos_m_x = 1e6f;
  ParticleFilter off_map(48, &map, fo, (uint32_t)seed + 4);
  // (d) the mixture the GMM thread fits drives the particle count of every update (particle_filter.cpp:151-158): with one
  // 10 x 12 px cluster 300 particles shrink to 3/4 + 10 per update; the EM input matrix comes straight off the device
  ParticleFilter adaptive(300, &map, fp, (uint32_t)seed + 5);
  adaptive.propagate(trans, 0.01f);
  auto gs = adaptive.gmmSamples();
  wr(dir + "gmm_samples.f64", gs.data(), gs.size());
  wr(dir + "gmm_states.bin", adaptive.states().data(), adaptive.states().size());
  std::vector<Vector3f> gm(1);
  std::vector<ParticleFilter::Matrix3f> gcov(1, ParticleFilter::Matrix3f{100, 0, 0, 0, 144, 0, 0, 0, 1});
  adaptive.setGMM(gm, gcov);
  adaptive.update(top_down, geo, fmeta[0]);
  const int n_adapt1 = adaptive.numParticles(), n_states1 = (int)adaptive.states().size();
  adaptive.update(top_down, geo, fmeta[0]);
  const int n_adapt2 = adaptive.numParticles(), n_states2 = (int)adaptive.states().size();
  const float misc[9] = {sc_fixed, sc_free, sc_frozen, (float)metric.numParticles(), (float)off_map.numParticles(),
                         (float)n_adapt1, (float)n_states1, (float)n_adapt2, (float)n_states2};
  wr(dir + "misc.f32", misc, 9);
  // ---- the vector-map constructor path and the raster cache (top_down_map.cpp:22-31, :197-224): polygons -> binary class
  // maps -> distance fields on the device; class<i>.png written, then a second map object starts from those files alone.
  // (one device context per process: from here on the device holds THIS map)
  {
    auto pv = rd<float>(dir + "polys.f32");
    auto ps = rd<int32_t>(dir + "poly_start.i32");
    auto pc = rd<int32_t>(dir + "poly_class.i32");
    auto vm = rd<int32_t>(dir + "vec_meta.i32");                                 // svg width, height, then the exclusive classes
    std::vector<std::vector<std::vector<Vector2f>>> poly(C);
    for (size_t k = 0; k < pc.size(); k++) {
      std::vector<Vector2f> path;
      for (int i = ps[k]; i < ps[k + 1]; i++) { Vector2f v; v.x = pv[2 * i]; v.y = pv[2 * i + 1]; path.push_back(v); }
      poly[pc[k]].push_back(path);
    }
    std::vector<int> excl(vm.begin() + 2, vm.end());
    TopDownMapPolar vmap(mp);
    vmap.samplePtsPolar(n_theta, n_r, (float)(2 * M_PI / n_theta));
    vmap.setVectorMap(poly, Vector2i{vm[0], vm[1]}, excl);
    if (!vmap.haveMap()) return 7;
    vmap.saveRasterizedMaps(dir + "vec_raster_cache");
    std::vector<ArrayXXf> loc(C, ArrayXXf(n_theta, n_r));
    ArrayXXc lmask(n_theta, n_r);
    Vector2f vc; vc.x = vm[0] * 0.45f; vc.y = vm[1] * 0.55f;
    vmap.getLocalMap(vc, 1.5f, 2.0f, loc, lmask);
    for (int c = 0; c < C; c++) wr(dir + "vec_local" + std::to_string(c) + ".f32", loc[c].data(), (size_t)n_theta * n_r);
    std::vector<int> at; vmap.getClassesAtPoint(vc, at);
    std::vector<int32_t> at32(at.begin(), at.end()); at32.push_back(-1);
    wr(dir + "vec_classes_at.i32", at32.data(), at32.size());
    TopDownMapPolar rmap(mp);
    rmap.samplePtsPolar(n_theta, n_r, (float)(2 * M_PI / n_theta));
    if (rmap.loadRasterizedMaps(dir + "no_such_cache")) return 8;
    if (!rmap.loadRasterizedMaps(dir + "vec_raster_cache") || !rmap.haveMap()) return 9;
    rmap.getLocalMap(vc, 1.5f, 2.0f, loc, lmask);
    for (int c = 0; c < C; c++) wr(dir + "ras_local" + std::to_string(c) + ".f32", loc[c].data(), (size_t)n_theta * n_r);
    wr(dir + "ras_mask.u8", lmask.data(), (size_t)n_theta * n_r);
  }
  std::cout << "host_demo ok: " << filter.numParticles() << " particles, mean (" << mean[0] << ", " << mean[1] << ", " << mean[2] << ")\n";
  return 0;
}
