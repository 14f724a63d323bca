// C++ parity tests through include/aboria_b200/Aboria.h (host compiler only,
// linked with libabr.so).  Ports of the reference's own tests for this path:
//   test_sparse_operator   /root/reference/tests/operators.h:810-933
//   test_documentation     /root/reference/tests/operators.h:121-311 (sparse part)
//   test_block_operator    /root/reference/tests/operators.h:963-1100 (with sparse blocks)
//   test_id_search         /root/reference/tests/id_search.h:64-99
// plus the container semantics of init_neighbour_search
// (src/NeighbourSearchBase.h:185-238, src/Particles.h:694-724).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "aboria_b200/Aboria.h"

using namespace Aboria;

static int failures = 0;
#define TS_ASSERT_EQUALS(a, b)                                                          \
  do {                                                                                  \
    if (!((a) == (b))) {                                                                \
      std::printf("FAIL %s:%d  %s == %s\n", __FILE__, __LINE__, #a, #b);               \
      ++failures;                                                                       \
    }                                                                                   \
  } while (0)
#define TS_ASSERT(a) TS_ASSERT_EQUALS(!!(a), true)

// Eigen::VectorXd stand-in: anything with (size) ctor, data(), size(), operator[]
typedef std::vector<double> vector_type;

static void test_sparse_operator() {
  ABORIA_VARIABLE(scalar1, double, "scalar1")
  ABORIA_VARIABLE(scalar2, double, "scalar2")
  typedef Particles<std::tuple<scalar1, scalar2>> ParticlesType;
  typedef position_d<3> position;
  ParticlesType particles;

  double diameter = 0.1;
  vdouble3 min = vdouble3::Constant(-1);
  vdouble3 max = vdouble3::Constant(1);
  vbool3 periodic = vbool3::Constant(false);

  double s_init1 = 1.0;
  double s_init2 = 2.0;
  ParticlesType::value_type p;
  get<position>(p) = vdouble3(0, 0, 0);
  get<scalar1>(p) = s_init1;
  get<scalar2>(p) = s_init2;
  particles.push_back(p);
  get<position>(p) = vdouble3(diameter * 0.9, 0, 0);
  particles.push_back(p);
  get<position>(p) = vdouble3(diameter * 1.8, 0, 0);
  particles.push_back(p);
  const size_t n = 3;

  particles.init_neighbour_search(min, max, periodic);

  //      3  3  0
  // C =  3  3  3
  //      0  3  3
  auto C = create_sparse_operator(particles, particles, diameter, kernels::const_sum<scalar1, scalar2>());
  vector_type v = {1, 2, 3};
  vector_type ans = C * v;
  for (size_t i = 0; i < n; i++) {
    double sum = 0;
    for (size_t j = 0; j < n; j++) {
      const size_t idi = get<id>(particles)[i], idj = get<id>(particles)[j];
      if ((idi == 0 && idj == 2) || (idi == 2 && idj == 0)) {
        sum += 0;
      } else {
        sum += (s_init1 + s_init2) * v[j];
      }
    }
    TS_ASSERT_EQUALS(ans[i], sum);
  }
  TS_ASSERT_EQUALS(ans[0], 9.0);
  TS_ASSERT_EQUALS(ans[1], 18.0);
  TS_ASSERT_EQUALS(ans[2], 15.0);

  //       3  3  0
  //      -1 -1  0
  // C2 =  3  3  3   (2x1 blocks)
  auto C2 = create_sparse_operator(particles, particles, diameter, kernels::const_sum_diff<scalar1, scalar2>());
  ans = C2 * v;
  std::vector<double> check = {9, -3, 18, -7, 15, -5};
  for (size_t i = 0; i < n; i++) TS_ASSERT_EQUALS(ans[i], check[i]); // the reference compares i < n only
  TS_ASSERT_EQUALS(ans.size(), 6u);
  TS_ASSERT_EQUALS(ans[4], 15.0);
  TS_ASSERT_EQUALS(ans[5], -5.0);

  // C.coeff(i, j) equals the assembled matrix (tests/operators.h:873-897): 7 non-zeros, all 3
  struct triplet {
    size_t r, c;
    double v;
    triplet(size_t r_, size_t c_, double v_) : r(r_), c(c_), v(v_) {}
  };
  std::vector<triplet> trip;
  C.assemble(trip);
  TS_ASSERT_EQUALS(trip.size(), 7u);
  double dense[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (const triplet &t : trip) {
    dense[t.r][t.c] = t.v;
    TS_ASSERT_EQUALS(t.v, C.coeff(t.r, t.c));
  }
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < n; j++) TS_ASSERT_EQUALS(dense[i][j], C.coeff(i, j));
  std::vector<triplet> trip2;
  C2.assemble(trip2);
  TS_ASSERT_EQUALS(trip2.size(), 14u); // tests/operators.h:938
  for (const triplet &t : trip2) TS_ASSERT_EQUALS(t.v, C2.coeff(t.r, t.c));

  // evaluate accumulates (lhs += K rhs)
  vector_type acc = {1, 1, 1};
  C.evaluate(acc, v);
  TS_ASSERT_EQUALS(acc[0], 10.0);
  TS_ASSERT_EQUALS(acc[1], 19.0);
  TS_ASSERT_EQUALS(acc[2], 16.0);
}

static void test_documentation() {
  const size_t N = 100;
  const double epsilon = 0.1;
  ABORIA_VARIABLE(a, double, "a");
  typedef Particles<std::tuple<a>> particle_type;
  typedef particle_type::position position;
  particle_type particles(N);
  std::default_random_engine gen;
  std::uniform_real_distribution<double> uniform(0, 1);
  for (size_t i = 0; i < N; ++i) {
    get<position>(particles)[i] = vdouble3(uniform(gen), uniform(gen), uniform(gen));
    get<a>(particles)[i] = uniform(gen);
  }
  const double r = 0.1;
  vdouble3 min = vdouble3::Constant(0);
  vdouble3 max = vdouble3::Constant(1);
  vbool3 periodic = vbool3::Constant(false);
  particles.init_neighbour_search(min, max, periodic);
  TS_ASSERT_EQUALS(particles.size(), N);
  auto K_s = create_sparse_operator(particles, particles, r, kernels::inv_dist_aa<a>(epsilon));
  vector_type b(N);
  for (size_t i = 0; i < N; ++i) b[i] = double(i) / (N - 1); // LinSpaced(N, 0, 1)
  vector_type c_3 = K_s * b;
  // c_4 = assembled sparse matrix * b (host loop over all pairs with the same predicate)
  double err2 = 0, nrm2 = 0;
  for (size_t i = 0; i < N; ++i) {
    double sum = 0;
    for (size_t j = 0; j < N; ++j) {
      const vdouble3 dx = get<position>(particles)[j] - get<position>(particles)[i];
      if (dx.squaredNorm() <= r * r) sum += (get<a>(particles)[i] * get<a>(particles)[j]) / (dx.norm() + epsilon) * b[j];
    }
    err2 += (sum - c_3[i]) * (sum - c_3[i]);
    nrm2 += sum * sum;
  }
  TS_ASSERT(std::sqrt(err2 / nrm2) <= 1e-12);
}

static void test_container_semantics() {
  // periodic wrap, kill outside a non-periodic dimension, reorder of every column
  ABORIA_VARIABLE(tag, double, "tag")
  typedef Particles<std::tuple<tag>, 2> P;
  typedef position_d<2> position;
  P particles;
  const double xs[6][2] = {{1.25, 0.5}, {-0.25, 0.5}, {0.5, 1.5}, {0.5, NAN}, {0.1, 0.2}, {3.75, 0.999}};
  for (int i = 0; i < 6; ++i) {
    P::value_type p;
    get<position>(p) = vdouble2(xs[i][0], xs[i][1]);
    get<tag>(p) = 100.0 + i;
    particles.push_back(p);
  }
  particles.init_neighbour_search(vdouble2(0, 0), vdouble2(1, 1), vbool2(true, false), 1.0);
  TS_ASSERT_EQUALS(particles.size(), 4u);
  for (size_t k = 0; k < particles.size(); ++k) {
    const size_t idk = get<id>(particles)[k];
    TS_ASSERT(idk == 0 || idk == 1 || idk == 4 || idk == 5);
    TS_ASSERT_EQUALS(get<tag>(particles)[k], 100.0 + idk); // columns travel together
    TS_ASSERT_EQUALS(get<alive>(particles)[k], 1);
    const double x = get<position>(particles)[k][0];
    if (idk == 0) TS_ASSERT_EQUALS(x, 0.25);
    if (idk == 1) TS_ASSERT_EQUALS(x, 0.75);
    if (idk == 5) TS_ASSERT_EQUALS(x, 0.75);
  }
}

static void test_block_operator() {
  // tests/operators.h:963-1100 builds [[A, B], [C, 0]] from dense blocks; dense kernels are
  // outside this path, so the same arrangement is made of sparse blocks
  ABORIA_VARIABLE(scalar1, double, "scalar1")
  ABORIA_VARIABLE(scalar2, double, "scalar2")
  typedef Particles<std::tuple<scalar1, scalar2>> ParticlesType;
  typedef position_d<3> position;
  ParticlesType particles, augment;
  const double diameter = 0.1;
  ParticlesType::value_type p;
  for (int i = 0; i < 3; ++i) {
    get<position>(p) = vdouble3(diameter * 0.9 * i, 0, 0);
    get<scalar1>(p) = 1 + i;
    get<scalar2>(p) = 0.1 * (1 + i);
    particles.push_back(p);
  }
  get<scalar1>(p) = 0;
  get<scalar2>(p) = 0.5;
  get<position>(p) = vdouble3(-diameter * 0.5, 0, 0);
  augment.push_back(p);
  particles.init_neighbour_search(vdouble3::Constant(-1), vdouble3::Constant(1), vbool3::Constant(false));
  augment.init_neighbour_search(vdouble3::Constant(-1), vdouble3::Constant(1), vbool3::Constant(false));
  const size_t n = 3;
  auto A = create_sparse_operator(particles, particles, diameter, kernels::const_sum<scalar1, scalar2>());
  auto B = create_sparse_operator(particles, augment, diameter, kernels::const_sum<scalar1, scalar2>());
  auto C = create_sparse_operator(augment, particles, diameter, kernels::const_sum<scalar2, scalar1>());
  auto Zero = create_zero_operator(augment, augment);
  auto Full = create_block_operator<2, 2>(A, B, C, Zero);
  TS_ASSERT_EQUALS(Full.rows(), n + 1);
  TS_ASSERT_EQUALS(Full.cols(), n + 1);
  vector_type v = {1, 2, 3, 4};
  vector_type ans = Full * v;
  // dense composition through coeff
  for (size_t i = 0; i < n + 1; ++i) {
    double sum = 0;
    for (size_t j = 0; j < n + 1; ++j) sum += Full.coeff(i, j) * v[j];
    TS_ASSERT(std::fabs(sum - ans[i]) <= 1e-14 * (1 + std::fabs(sum)));
  }
  TS_ASSERT_EQUALS(Full.coeff(n, n), 0.0);
  // particle at -0.05 is within 0.1 of x = 0 only
  size_t nz_last_row = 0;
  for (size_t j = 0; j < n; ++j) nz_last_row += Full.coeff(n, j) != 0.0;
  TS_ASSERT_EQUALS(nz_last_row, 1u);
  struct triplet {
    size_t r, c;
    double v;
    triplet(size_t r_, size_t c_, double v_) : r(r_), c(c_), v(v_) {}
  };
  std::vector<triplet> trip;
  Full.assemble(trip);
  TS_ASSERT_EQUALS(trip.size(), 7u + 1u + 1u);
  for (const triplet &t : trip) TS_ASSERT_EQUALS(t.v, Full.coeff(t.r, t.c));
}

static void test_id_search() {
  // tests/id_search.h:64-99
  const size_t N = 100;
  typedef Particles<> particle_type;
  particle_type particles(N);
  std::default_random_engine g;
  auto &ids = get<id>(particles);
  std::shuffle(ids.begin(), ids.end(), g);
  particles.init_id_search();
  const size_t id_2 = particles.get_query().find(2);
  TS_ASSERT(id_2 < N);
  TS_ASSERT_EQUALS(get<id>(particles)[id_2], 2u);
  const size_t id_2N = particles.get_query().find(2 * N);
  TS_ASSERT_EQUALS(id_2N, particles.size()); // "end of the particle vector"
  // the map follows the reorder of init_neighbour_search
  std::uniform_real_distribution<double> uniform(0, 1);
  for (size_t i = 0; i < N; ++i) get<particle_type::position>(particles)[i] = vdouble3(uniform(g), uniform(g), uniform(g));
  particles.init_neighbour_search(vdouble3::Constant(0), vdouble3::Constant(1), vbool3::Constant(true));
  for (size_t want = 0; want < N; want += 7) TS_ASSERT_EQUALS(get<id>(particles)[particles.get_query().find(want)], want);
}

static void test_md_force_block_kernel() {
  // tests/md.h:166-174: sum over neighbours of -k (diameter / r - 1) dx as a D x 1 block operator times ones
  typedef Particles<std::tuple<>, 2> P;
  typedef position_d<2> position;
  const size_t N = 400;
  const double diameter = 0.08, k = 10.0;
  P particles(N);
  std::default_random_engine gen(5);
  std::uniform_real_distribution<double> uniform(0, 1);
  for (size_t i = 0; i < N; ++i) get<position>(particles)[i] = vdouble2(uniform(gen), uniform(gen));
  particles.init_neighbour_search(vdouble2(0, 0), vdouble2(1, 1), vbool2(false, false));
  auto F = create_sparse_operator(particles, particles, diameter, kernels::linear_spring<2>(k, diameter));
  vector_type ones(N, 1.0);
  vector_type f = F * ones;
  TS_ASSERT_EQUALS(f.size(), 2 * N);
  double err2 = 0, nrm2 = 0;
  for (size_t i = 0; i < N; ++i) {
    double sx = 0, sy = 0;
    for (size_t j = 0; j < N; ++j) {
      const vdouble2 dx = get<position>(particles)[j] - get<position>(particles)[i];
      const double r = dx.norm();
      if (dx.squaredNorm() <= diameter * diameter && r != 0) {
        sx += -k * (diameter / r - 1.0) * dx[0];
        sy += -k * (diameter / r - 1.0) * dx[1];
      }
    }
    err2 += (sx - f[2 * i]) * (sx - f[2 * i]) + (sy - f[2 * i + 1]) * (sy - f[2 * i + 1]);
    nrm2 += sx * sx + sy * sy;
  }
  TS_ASSERT(nrm2 > 0);
  TS_ASSERT(std::sqrt(err2 / nrm2) <= 1e-12);
}

// A row set WITHOUT a neighbour search whose functor reads a row column (the row
// columns must be shipped, not null / stale), the FRadius overload
// (/root/reference/src/Operators.h:478-489), push_back with update_neighbour_search
// (src/Particles.h:264-297) and device-resident vectors for solver loops.
static void test_rows_without_search_fradius_pushback_devicevector() {
  ABORIA_VARIABLE(scalar1, double, "scalar1")
  ABORIA_VARIABLE(scalar2, double, "scalar2")
  typedef Particles<std::tuple<scalar1, scalar2>> ParticlesType;
  typedef position_d<3> position;
  ParticlesType cols, rows;
  const double diameter = 0.1;
  ParticlesType::value_type p;
  for (int i = 0; i < 3; ++i) {
    get<position>(p) = vdouble3(diameter * 0.9 * i, 0, 0);
    get<scalar1>(p) = 1.0;
    get<scalar2>(p) = 2.0 + i; // 2, 3, 4
    cols.push_back(p);
  }
  cols.init_neighbour_search(vdouble3::Constant(-1), vdouble3::Constant(1), vbool3::Constant(false));
  // two test points, never given a search structure; their scalar1 differs per row
  get<position>(p) = vdouble3(0.01, 0, 0);
  get<scalar1>(p) = 10.0;
  rows.push_back(p);
  get<position>(p) = vdouble3(0.17, 0, 0);
  get<scalar1>(p) = 20.0;
  rows.push_back(p);
  auto G = create_sparse_operator(rows, cols, diameter, kernels::const_sum<scalar1, scalar2>());
  vector_type v(3, 1.0);
  vector_type y = G * v;
  // row 0 (x=0.01) reaches columns at 0 and 0.09: (10+2) + (10+3) = 25; row 1 (x=0.17) reaches 0.09 and 0.18: (20+3) + (20+4) = 47
  TS_ASSERT_EQUALS(y.size(), (size_t)2);
  TS_ASSERT_EQUALS(y[0], 25.0);
  TS_ASSERT_EQUALS(y[1], 47.0);
  // the reference reads LIVE row values: edit the host column, apply again
  get<scalar1>(rows)[0] = 100.0;
  y = G * v;
  TS_ASSERT_EQUALS(y[0], 205.0);
  TS_ASSERT_EQUALS(G.coeff(1, 2), 24.0);
  TS_ASSERT_EQUALS(G.coeff(1, 0), 0.0);

  // FRadius: the radius is a function of the row particle
  auto Gr = create_sparse_operator(rows, cols, [&](const ParticlesType::value_type &a) { return get<scalar1>(a) > 50.0 ? 0.05 : 0.2; },
                                   kernels::const_sum<scalar1, scalar2>());
  y = Gr * v;
  TS_ASSERT_EQUALS(y[0], 102.0);                                   // radius 0.05: only the column at 0
  TS_ASSERT_EQUALS(y[1], (20.0 + 2) + (20.0 + 3) + (20.0 + 4));    // radius 0.2: all three

  // push_back on a searchable container refreshes the ordered structure (the default) ...
  get<position>(p) = vdouble3(-0.05, 0, 0);
  get<scalar1>(p) = 1.0;
  get<scalar2>(p) = 7.0;
  cols.push_back(p);
  TS_ASSERT(cols.searchable());
  TS_ASSERT_EQUALS(cols.size(), (size_t)4);
  {
    bool found = false; // the reorder kept the new particle (4 particles, one bucket: stable order, it comes last)
    for (size_t i = 0; i < cols.size(); ++i) found |= get<position>(cols)[i][0] == -0.05;
    TS_ASSERT(found);
  }
  vector_type v4(4, 1.0);
  auto G4 = create_sparse_operator(rows, cols, diameter, kernels::const_sum<scalar1, scalar2>());
  y = G4 * v4;
  TS_ASSERT_EQUALS(y[0], 205.0 + 107.0);
  // ... unless the caller defers it
  cols.push_back(p, false);
  TS_ASSERT(!cols.searchable());
  cols.update_positions();
  TS_ASSERT(cols.searchable());

  // device-resident vectors: y = K b without host copies, twice on the same buffers
  auto C = create_sparse_operator(cols, cols, diameter, kernels::const_sum<scalar1, scalar2>());
  vector_type b5(cols.size(), 1.0), y_host = C * b5;
  DeviceVector bd(cols.handle(), b5), yd(cols.handle(), cols.size());
  yd.set_zero();
  C.evaluate(yd, bd);
  vector_type y_dev(cols.size());
  yd.download(y_dev);
  for (size_t i = 0; i < cols.size(); ++i) TS_ASSERT_EQUALS(y_dev[i], y_host[i]);
  C.evaluate(yd, bd); // accumulates
  yd.download(y_dev);
  for (size_t i = 0; i < cols.size(); ++i) TS_ASSERT_EQUALS(y_dev[i], 2 * y_host[i]);
  DeviceVector y2 = C * bd;
  y2.download(y_dev);
  for (size_t i = 0; i < cols.size(); ++i) TS_ASSERT_EQUALS(y_dev[i], y_host[i]);
}

// Level 3 names (src/Symbolic.h:269-444): r[a] = sum(b, summand) with AccumulateWithinDistance<std::plus<double>>
static void test_accumulate_within_distance() {
  ABORIA_VARIABLE(scalar1, double, "scalar1")
  ABORIA_VARIABLE(scalar2, double, "scalar2")
  ABORIA_VARIABLE(total, double, "total")
  typedef Particles<std::tuple<scalar1, scalar2, total>> ParticlesType;
  typedef position_d<3> position;
  ParticlesType particles;
  const double diameter = 0.1;
  ParticlesType::value_type p;
  for (int i = 0; i < 3; ++i) {
    get<position>(p) = vdouble3(diameter * 0.9 * i, 0, 0);
    get<scalar1>(p) = 1.0;
    get<scalar2>(p) = 2.0;
    get<total>(p) = -1.0;
    particles.push_back(p);
  }
  particles.init_neighbour_search(vdouble3::Constant(-1), vdouble3::Constant(1), vbool3::Constant(false));
  Symbol<total> t;
  Label<0, ParticlesType> a(particles);
  Label<1, ParticlesType> b(particles);
  auto dx = create_dx(a, b);
  (void)dx;
  AccumulateWithinDistance<std::plus<double>> sum(diameter);
  t[a] = sum(b, kernels::const_sum<scalar1, scalar2>()); // sum over neighbours of (scalar1_a + scalar2_b): rows of C = [3 3 0; 3 3 3; 0 3 3]
  TS_ASSERT_EQUALS(get<total>(particles)[0], 6.0);
  TS_ASSERT_EQUALS(get<total>(particles)[1], 9.0);
  TS_ASSERT_EQUALS(get<total>(particles)[2], 6.0);
  sum.set_init(0.5);
  sum.set_max_distance(2 * diameter);
  t[a] = sum(b, kernels::const_sum<scalar1, scalar2>());
  TS_ASSERT_EQUALS(get<total>(particles)[0], 9.5);
}

int main() {
  test_accumulate_within_distance();
  test_rows_without_search_fradius_pushback_devicevector();
  test_md_force_block_kernel();
  test_sparse_operator();
  test_block_operator();
  test_id_search();
  test_documentation();
  test_container_semantics();
  if (failures) {
    std::printf("%d failure(s)\n", failures);
    return 1;
  }
  std::printf("all shim tests passed\n");
  return 0;
}
