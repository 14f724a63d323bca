// Aboria.h — C++ mirror of the reference interface for the accelerated path.
//
// Keeps the reference's names and call shapes (/root/reference/src):
//   ABORIA_VARIABLE(name, type, "desc")                     Variable.h:75-79
//   Particles<std::tuple<vars...>, D>                        Particles.h:108-112
//     push_back / size / operator[] / get<var>(particles)    Particles.h, Get.h:1110-1150
//     init_neighbour_search(low, high, periodic, n_leaf)     Particles.h:445-455
//     update_positions()                                     Particles.h:526-531, 694-724
//   create_sparse_operator(rows, cols, radius, f)            Operators.h:478-516
//   K * b                                                    Operators.h:153 -> Kernels.h:720-751
//   K.coeff(i, j), K.assemble(triplets)                      Operators.h:149-151, Kernels.h:653-685
//   create_zero_operator, create_block_operator<NI,NJ>       Operators.h:531-548
//   init_id_search(), get_query().find(id)                   NeighbourSearchBase.h:294-298, CellListOrdered.h:379-388
// on top of the C-ABI in abr.h.  Plain C++14, no CUDA headers needed: the user
// TU is compiled by the host compiler and linked with libabr.so.
//
// What differs from the reference, and why:
//   * the kernel function of create_sparse_operator is a *kernel descriptor*
//     (Aboria::kernels::const_sum<s1,s2>(), inv_dist(eps), wendland_c2(h), ...)
//     or, in a .cu file, a device functor (aboria_b200/device_kernel.cuh):
//     a host lambda cannot run on the GPU (the reference itself marks the
//     kernel API host-only, tests/parallel.h:60-62).
//   * vectors are any type with data()/size() and a (size) constructor —
//     Eigen::VectorXd fits, and so does std::vector<double>; Eigen itself is not
//     required.
//   * errors follow the reference's CHECK (message to std::cerr + SIGTRAP,
//     src/Log.h:57-62).
// Columns live in host std::vectors exactly like Particles<..., std::vector>;
// update_positions ships them to the GPU, runs the build + reorder there and
// copies the reordered columns back (the reference mutates the container the
// same way, NeighbourSearchBase.h:185-238 and Particles.h:694-724); device
// copies stay resident for the products.
#ifndef ABORIA_B200_ABORIA_H_
#define ABORIA_B200_ABORIA_H_

#include <array>
#include <cmath>
#include <csignal>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <string>
#include <tuple>
#include <type_traits>
#include <functional>
#include <vector>

#include "abr.h"

#define ABR_CHECK(cond, message)                                                                     \
  if (!(cond)) {                                                                                     \
    std::cerr << "Aboria(b200) CHECK failed: " << message << " (" << __FILE__ << ":" << __LINE__ << ")" \
              << std::endl;                                                                          \
    std::raise(SIGTRAP);                                                                             \
  }

namespace Aboria {

// ---- Vector<T,N> (src/Vector.h:100), the POD layout only -------------------
template <typename T, unsigned int N> struct Vector {
  T mem[N];
  Vector() {
    for (unsigned i = 0; i < N; ++i) mem[i] = T();
  }
  template <typename... Args, typename = typename std::enable_if<sizeof...(Args) == N && (N > 1)>::type>
  Vector(Args... args) : mem{static_cast<T>(args)...} {}
  static Vector Constant(const T &c) {
    Vector v;
    for (unsigned i = 0; i < N; ++i) v.mem[i] = c;
    return v;
  }
  T &operator[](unsigned i) { return mem[i]; }
  const T &operator[](unsigned i) const { return mem[i]; }
  double squaredNorm() const {
    double ret = 0;
    for (unsigned i = 0; i < N; ++i) ret += mem[i] * mem[i];
    return ret;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
};
template <typename T, unsigned int N> Vector<double, N> operator-(const Vector<T, N> &a, const Vector<T, N> &b) {
  Vector<double, N> r;
  for (unsigned i = 0; i < N; ++i) r[i] = a[i] - b[i];
  return r;
}
typedef Vector<double, 1> vdouble1;
typedef Vector<double, 2> vdouble2;
typedef Vector<double, 3> vdouble3;
typedef Vector<bool, 1> vbool1;
typedef Vector<bool, 2> vbool2;
typedef Vector<bool, 3> vbool3;

// ---- variables (src/Variable.h) ---------------------------------------------
#define ABORIA_VARIABLE(NAME, DATA_TYPE, NAME_STRING)  \
  struct NAME {                                        \
    typedef DATA_TYPE value_type;                      \
    static const char *name() { return NAME_STRING; } \
  };

template <unsigned int D> struct position_d {
  typedef Vector<double, D> value_type;
  static const char *name() { return "position"; }
};
struct id {
  typedef size_t value_type;
  static const char *name() { return "id"; }
};
struct alive {
  typedef uint8_t value_type;
  static const char *name() { return "alive"; }
};

namespace detail {
template <typename T, typename Tuple> struct index_of;
template <typename T, typename... Ts> struct index_of<T, std::tuple<T, Ts...>> { static const size_t value = 0; };
template <typename T, typename U, typename... Ts> struct index_of<T, std::tuple<U, Ts...>> {
  static const size_t value = 1 + index_of<T, std::tuple<Ts...>>::value;
};
template <typename... Vars> struct columns_of { typedef std::tuple<std::vector<typename Vars::value_type>...> type; };

inline void check_rc(abr_handle h, int rc, const char *what) {
  ABR_CHECK(rc == 0, what << ": " << abr_last_error_string(h));
}
} // namespace detail

// one particle (value_type of the reference container)
template <unsigned int D, typename... UserVars> struct particle_value {
  typedef std::tuple<position_d<D>, id, alive, UserVars...> variables;
  std::tuple<Vector<double, D>, size_t, uint8_t, typename UserVars::value_type...> v;
  particle_value() { std::get<2>(v) = uint8_t(true); }
};
// get<variable>(particle)
template <typename Var, unsigned int D, typename... UserVars>
typename Var::value_type &get(particle_value<D, UserVars...> &p) {
  return std::get<detail::index_of<Var, typename particle_value<D, UserVars...>::variables>::value>(p.v);
}

template <typename Var, unsigned int D, typename... UserVars>
const typename Var::value_type &get(const particle_value<D, UserVars...> &p) {
  return std::get<detail::index_of<Var, typename particle_value<D, UserVars...>::variables>::value>(p.v);
}

// ---- Particles ---------------------------------------------------------------
template <typename VAR = std::tuple<>, unsigned int DomainD = 3> class Particles;

template <typename... UserVars, unsigned int DomainD> class Particles<std::tuple<UserVars...>, DomainD> {
public:
  static const unsigned int dimension = DomainD;
  typedef position_d<DomainD> position;
  typedef Vector<double, DomainD> double_d;
  typedef Vector<bool, DomainD> bool_d;
  typedef std::tuple<position, id, alive, UserVars...> variables;
  typedef std::tuple<std::vector<double_d>, std::vector<size_t>, std::vector<uint8_t>,
                     std::vector<typename UserVars::value_type>...>
      data_type;
  static const size_t n_columns = 3 + sizeof...(UserVars);

  typedef particle_value<DomainD, UserVars...> value_type;

  Particles() : next_id_(0), searchable_(false), id_map_(false) { open(); }
  explicit Particles(size_t n) : next_id_(0), searchable_(false), id_map_(false) {
    open();
    resize(n);
  }
  Particles(const Particles &) = delete;
  Particles &operator=(const Particles &) = delete;
  ~Particles() {
    free_device();
    if (h_) abr_destroy(h_);
  }

  size_t size() const { return std::get<0>(data_).size(); }
  void resize(size_t n) {
    const size_t old = size();
    resize_all(n, std::make_index_sequence<n_columns>());
    for (size_t i = old; i < n; ++i) {
      std::get<1>(data_)[i] = next_id_++;
      std::get<2>(data_)[i] = uint8_t(true);
    }
  }
  // src/Particles.h:264-297: with a neighbour search in place the ordered structure is rebuilt (the
  // reference's update_positions(begin(), end()) for search.ordered()) unless the caller asks not to —
  // then the container is no longer searchable until the next update_positions
  void push_back(const value_type &p, bool update_neighbour_search = true) {
    push_all(p, std::make_index_sequence<n_columns>());
    std::get<1>(data_).back() = next_id_++;
    std::get<2>(data_).back() = uint8_t(true);
    if (searchable_ && update_neighbour_search) {
      update_positions();
    } else {
      searchable_ = false;
    }
  }
  // the particle at position i, by value (the reference hands functors a const reference)
  value_type operator[](size_t i) const {
    value_type p;
    get_all(p, i, std::make_index_sequence<n_columns>());
    return p;
  }

  template <typename Var> std::vector<typename Var::value_type> &column() {
    return std::get<detail::index_of<Var, variables>::value>(data_);
  }
  template <typename Var> const std::vector<typename Var::value_type> &column() const {
    return std::get<detail::index_of<Var, variables>::value>(data_);
  }

  // src/Particles.h:445-455
  void init_neighbour_search(const double_d &low, const double_d &high, const bool_d &periodic,
                             const double n_particles_in_leaf = 10.0) {
    double lo[DomainD], hi[DomainD];
    uint8_t per[DomainD];
    for (unsigned d = 0; d < DomainD; ++d) {
      lo[d] = low[d];
      hi[d] = high[d];
      per[d] = periodic[d] ? 1 : 0;
    }
    detail::check_rc(h_, abr_domain_set(h_, (int)DomainD, lo, hi, per, n_particles_in_leaf), "init_neighbour_search");
    update_positions();
  }

  // src/Particles.h:526-531 (+ reorder :694-724): wrap/kill, ordered cell list,
  // reorder every column; dead particles are removed
  void update_positions() {
    const size_t n = size();
    free_device();
    alloc_device(n);
    upload_all(std::make_index_sequence<n_columns>());
    const void *src[n_columns];
    void *dst[n_columns];
    size_t eb[n_columns];
    for (size_t c = 0; c < n_columns; ++c) {
      src[c] = dev_[c];
      dst[c] = dev_other_[c];
      eb[c] = elem_bytes_[c];
    }
    int32_t *order = nullptr;
    detail::check_rc(h_, abr_malloc(h_, (void **)&order, (n + 1) * sizeof(int32_t)), "update_positions");
    size_t n_alive = 0;
    detail::check_rc(h_,
                     abr_update_positions(h_, static_cast<double *>(dev_[0]), static_cast<uint8_t *>(dev_[2]), n, (int)n_columns,
                                          src, dst, eb, order, &n_alive),
                     "update_positions");
    abr_free(h_, order);
    for (size_t c = 0; c < n_columns; ++c) std::swap(dev_[c], dev_other_[c]); // data.swap(other_data)
    resize_all(n_alive, std::make_index_sequence<n_columns>());
    download_all(std::make_index_sequence<n_columns>());
    n_device_ = n_alive;
    searchable_ = true;
    if (id_map_) update_id_map();
  }

  // find-by-id (src/NeighbourSearchBase.h:294-298, :440-486): an id -> position map sorted
  // by id, rebuilt by every update_positions like the reference's
  void init_id_search() {
    id_map_ = true;
    if (n_device_ != size() || !dev_[1]) { // no device copy yet (no neighbour search): ship the id column
      free_device();
      alloc_device(size());
      upload_all(std::make_index_sequence<n_columns>());
      n_device_ = size();
    }
    update_id_map();
  }
  // the part of CellListOrderedQuery this path uses (src/CellListOrdered.h:285-599)
  struct Query {
    const Particles *p;
    size_t number_of_particles() const { return p->size(); }
    // find(id) (src/CellListOrdered.h:379-388): the reference returns a pointer into the
    // particle set, `begin + n` when the id is absent; here: the index, n when absent
    size_t find(const size_t id_to_find) const {
      ABR_CHECK(p->id_map_, "init_id_search not called on this particle set");
      abr_handle h = p->h_;
      uint64_t *q = nullptr;
      detail::check_rc(h, abr_malloc(h, (void **)&q, 2 * sizeof(uint64_t)), "find");
      const uint64_t idv = id_to_find;
      uint64_t out = 0;
      detail::check_rc(h, abr_memcpy_h2d(h, q, &idv, sizeof(idv)), "find");
      detail::check_rc(h, abr_id_find(h, q, 1, q + 1), "find");
      detail::check_rc(h, abr_memcpy_d2h(h, &out, q + 1, sizeof(out)), "find");
      abr_free(h, q);
      return (size_t)out;
    }
  };
  Query get_query() const { return Query{this}; }

  bool searchable() const { return searchable_; }
  abr_handle handle() const { return h_; }
  // device pointer of a column (valid until the next update_positions)
  template <typename Var> const void *device_column() const { return dev_[detail::index_of<Var, variables>::value]; }
  const double *device_positions() const { return static_cast<const double *>(dev_[0]); }
  size_t device_size() const { return n_device_; }
  // Ship every column of this set to the device as it is on the host NOW (no search structure needed).
  // An operator whose row set is not its column set calls this before it binds row columns: the
  // reference reads live host values of the row particles (src/Kernels.h:737-749), so a row set that never
  // ran update_positions, or whose host columns were edited since, must not leave a null / stale pointer.
  void sync_all_to_device() const {
    if (n_device_ != size() || !dev_[0]) {
      free_device();
      alloc_device(size());
      n_device_ = size();
    }
    upload_all(std::make_index_sequence<n_columns>());
  }
  // copy a (possibly modified) host column to the device again
  template <typename Var> void sync_to_device() {
    const size_t c = detail::index_of<Var, variables>::value;
    ABR_CHECK(searchable_ && n_device_ == size(), "sync_to_device: call update_positions first");
    detail::check_rc(h_, abr_memcpy_h2d(h_, dev_[c], column<Var>().data(), elem_bytes_[c] * size()), "sync_to_device");
  }

private:
  void update_id_map() {
    detail::check_rc(h_, abr_id_map_build(h_, static_cast<const uint64_t *>(dev_[1]), n_device_), "init_id_search");
  }
  void open() {
    h_ = nullptr;
    int rc = abr_create(&h_, 0, nullptr);
    ABR_CHECK(rc == 0, "abr_create: " << abr_last_error_string(nullptr));
    for (size_t c = 0; c < n_columns; ++c) dev_[c] = dev_other_[c] = nullptr;
    set_elem_bytes(std::make_index_sequence<n_columns>());
    n_device_ = 0;
  }
  template <size_t... I> void set_elem_bytes(std::index_sequence<I...>) {
    size_t s[] = {sizeof(typename std::tuple_element<I, data_type>::type::value_type)...};
    for (size_t c = 0; c < n_columns; ++c) elem_bytes_[c] = s[c];
  }
  template <size_t... I> void resize_all(size_t n, std::index_sequence<I...>) {
    int dummy[] = {(std::get<I>(data_).resize(n), 0)...};
    (void)dummy;
  }
  template <size_t... I> void get_all(value_type &p, size_t i, std::index_sequence<I...>) const {
    int dummy[] = {(std::get<I>(p.v) = std::get<I>(data_)[i], 0)...};
    (void)dummy;
  }
  template <size_t... I> void push_all(const value_type &p, std::index_sequence<I...>) {
    int dummy[] = {(std::get<I>(data_).push_back(std::get<I>(p.v)), 0)...};
    (void)dummy;
  }
  template <size_t... I> void upload_all(std::index_sequence<I...>) const {
    int dummy[] = {(detail::check_rc(h_, abr_memcpy_h2d(h_, dev_[I], std::get<I>(data_).data(), elem_bytes_[I] * size()), "upload"), 0)...};
    (void)dummy;
  }
  template <size_t... I> void download_all(std::index_sequence<I...>) {
    int dummy[] = {(detail::check_rc(h_, abr_memcpy_d2h(h_, std::get<I>(data_).data(), dev_[I], elem_bytes_[I] * size()), "download"), 0)...};
    (void)dummy;
  }
  void alloc_device(size_t n) const {
    for (size_t c = 0; c < n_columns; ++c) {
      detail::check_rc(h_, abr_malloc(h_, &dev_[c], elem_bytes_[c] * (n + 1)), "alloc");
      detail::check_rc(h_, abr_malloc(h_, &dev_other_[c], elem_bytes_[c] * (n + 1)), "alloc");
    }
  }
  void free_device() const {
    for (size_t c = 0; c < n_columns; ++c) {
      if (dev_[c]) abr_free(h_, dev_[c]);
      if (dev_other_[c]) abr_free(h_, dev_other_[c]);
      dev_[c] = dev_other_[c] = nullptr;
    }
  }

  data_type data_;
  size_t next_id_;
  bool searchable_;
  bool id_map_;
  abr_handle h_;
  mutable void *dev_[n_columns], *dev_other_[n_columns]; // device copies: a cache of the host columns
  size_t elem_bytes_[n_columns];
  mutable size_t n_device_;
};

// get<variable>(particles)  (src/Get.h:1110-1150)
template <typename Var, typename P> auto get(P &particles) -> decltype(particles.template column<Var>()) {
  return particles.template column<Var>();
}

// ---- kernel descriptors --------------------------------------------------------
// Each stands for a lambda of the reference's tests; `bind` resolves variable
// columns to device pointers when the operator is applied.
namespace kernels {
struct desc_base {
  abr_kernel_desc d;
  desc_base() { std::memset(&d, 0, sizeof(d)); }
};
// get<S1>(a) + get<S2>(b)   (tests/operators.h:842-847)
template <typename S1, typename S2> struct const_sum : desc_base {
  const_sum() {
    d.kernel_id = ABR_K_CONST_SUM;
    d.block_rows = d.block_cols = 1;
  }
  template <typename R, typename Cc> void bind(const R &rows, const Cc &cols) {
    d.row_vars[0] = static_cast<const double *>(rows.template device_column<S1>());
    d.col_vars[0] = static_cast<const double *>(cols.template device_column<S2>());
  }
};
// 2x1 block (s1(a)+s2(b), s1(a)-s2(b))   (tests/operators.h:905-911)
template <typename S1, typename S2> struct const_sum_diff : desc_base {
  const_sum_diff() {
    d.kernel_id = ABR_K_CONST_SUM_DIFF;
    d.block_rows = 2;
    d.block_cols = 1;
  }
  template <typename R, typename Cc> void bind(const R &rows, const Cc &cols) {
    d.row_vars[0] = static_cast<const double *>(rows.template device_column<S1>());
    d.col_vars[0] = static_cast<const double *>(cols.template device_column<S2>());
  }
};
// 1/(|dx| + eps)
struct inv_dist : desc_base {
  explicit inv_dist(double eps) {
    d.kernel_id = ABR_K_INV_DIST;
    d.block_rows = d.block_cols = 1;
    d.params[0] = eps;
  }
  template <typename R, typename Cc> void bind(const R &, const Cc &) {}
};
// get<A>(i) get<A>(j) / (|dx| + eps)   (tests/operators.h:251-256)
template <typename A> struct inv_dist_aa : desc_base {
  explicit inv_dist_aa(double eps) {
    d.kernel_id = ABR_K_INV_DIST_AA;
    d.block_rows = d.block_cols = 1;
    d.params[0] = eps;
  }
  template <typename R, typename Cc> void bind(const R &rows, const Cc &cols) {
    d.row_vars[0] = static_cast<const double *>(rows.template device_column<A>());
    d.col_vars[0] = static_cast<const double *>(cols.template device_column<A>());
  }
};
// pow(2 - |dx|/h, 4) * (1 + 2|dx|/h)   (tests/rbf_interpolation.h:310-313)
struct wendland_c2 : desc_base {
  explicit wendland_c2(double h) {
    d.kernel_id = ABR_K_WENDLAND_C2;
    d.block_rows = d.block_cols = 1;
    d.params[0] = h;
  }
  template <typename R, typename Cc> void bind(const R &, const Cc &) {}
};
// D x 1 Lennard-Jones force 24 eps (2 (sigma/r)^12 - (sigma/r)^6) / r^2 dx   (tests/md.h pattern)
template <unsigned int D> struct lj_force : desc_base {
  lj_force(double sigma, double eps) {
    d.kernel_id = ABR_K_LJ_FORCE;
    d.block_rows = D;
    d.block_cols = 1;
    d.params[0] = sigma;
    d.params[1] = eps;
  }
  template <typename R, typename Cc> void bind(const R &, const Cc &) {}
};
// D x 1 linear spring -k (diameter/|dx| - 1) dx, 0 at |dx| = 0   (tests/md.h:166-174)
template <unsigned int D> struct linear_spring : desc_base {
  linear_spring(double k, double diameter) {
    d.kernel_id = ABR_K_LINEAR_SPRING;
    d.block_rows = D;
    d.block_cols = 1;
    d.params[0] = k;
    d.params[1] = diameter;
  }
  template <typename R, typename Cc> void bind(const R &, const Cc &) {}
};
// mass * W(|dx|, h)   (tests/sph.h:154-165)
struct sph_density : desc_base {
  sph_density(double h, double mass, double wcon) {
    d.kernel_id = ABR_K_SPH_DENSITY;
    d.block_rows = d.block_cols = 1;
    d.params[0] = h;
    d.params[1] = mass;
    d.params[2] = wcon;
  }
  template <typename R, typename Cc> void bind(const R &, const Cc &) {}
};
// D x 1: mass (get<P>(a) + get<P>(b)) F(|dx|, h) dx, P = pressure / density^2   (tests/sph.h:140-152, :333-339)
template <unsigned int D, typename P> struct sph_pressure : desc_base {
  sph_pressure(double h, double mass, double wcon) {
    d.kernel_id = ABR_K_SPH_PRESSURE;
    d.block_rows = D;
    d.block_cols = 1;
    d.params[0] = h;
    d.params[1] = mass;
    d.params[2] = wcon;
  }
  template <typename R, typename Cc> void bind(const R &rows, const Cc &cols) {
    d.row_vars[0] = static_cast<const double *>(rows.template device_column<P>());
    d.col_vars[0] = static_cast<const double *>(cols.template device_column<P>());
  }
};
} // namespace kernels

// ---- device-resident vector ---------------------------------------------------------
// b and y of `K * b` kept on the GPU across the iterations of a solver (the reference's
// Eigen::VectorXd lives on the host; shipping 2 x 256 MB over PCIe per product costs more than
// the product itself).  Owns its memory through the C-ABI; movable, not copyable.
class DeviceVector {
public:
  DeviceVector(abr_handle h, size_t n) : h_(h), n_(n), d_(nullptr) {
    detail::check_rc(h_, abr_malloc(h_, (void **)&d_, sizeof(double) * (n + 1)), "DeviceVector");
  }
  template <typename HostVector> DeviceVector(abr_handle h, const HostVector &v) : DeviceVector(h, (size_t)v.size()) { upload(v); }
  DeviceVector(DeviceVector &&o) noexcept : h_(o.h_), n_(o.n_), d_(o.d_) { o.d_ = nullptr; }
  DeviceVector(const DeviceVector &) = delete;
  DeviceVector &operator=(const DeviceVector &) = delete;
  ~DeviceVector() {
    if (d_) abr_free(h_, d_);
  }
  size_t size() const { return n_; }
  double *data() { return d_; }
  const double *data() const { return d_; }
  void set_zero() { detail::check_rc(h_, abr_memset(h_, d_, 0, sizeof(double) * n_), "DeviceVector"); }
  template <typename HostVector> void upload(const HostVector &v) {
    ABR_CHECK((size_t)v.size() == n_, "DeviceVector: size mismatch");
    detail::check_rc(h_, abr_memcpy_h2d(h_, d_, v.data(), sizeof(double) * n_), "DeviceVector");
  }
  template <typename HostVector> void download(HostVector &v) const {
    ABR_CHECK((size_t)v.size() == n_, "DeviceVector: size mismatch");
    detail::check_rc(h_, abr_memcpy_d2h(h_, v.data(), d_, sizeof(double) * n_), "DeviceVector");
  }

private:
  abr_handle h_;
  size_t n_;
  double *d_;
};

// ---- sparse operator ------------------------------------------------------------
// MatrixReplacement<1,1,tuple<KernelSparseConst>> (src/Operators.h:75-291);
// stores REFERENCES to the particle sets like the reference (src/Kernels.h:133-134)
template <typename RowParticles, typename ColParticles, typename KernelDesc> class SparseOperator {
public:
  SparseOperator(const RowParticles &rows, const ColParticles &cols, double radius, const KernelDesc &k)
      : rows_(rows), cols_(cols), radius_(radius), k_(k) {}
  size_t rows() const { return rows_.size() * k_.d.block_rows; }
  size_t cols() const { return cols_.size() * k_.d.block_cols; }

  // lhs += K rhs   (KernelSparse::evaluate, src/Kernels.h:720-751)
  template <typename LHS, typename RHS> void evaluate(LHS &lhs, const RHS &rhs) const {
    ABR_CHECK((size_t)lhs.size() == rows(), "lhs vector has incompatible size");
    ABR_CHECK((size_t)rhs.size() == cols(), "rhs vector has incompatible size");
    ABR_CHECK(cols_.searchable(), "column particles have no neighbour search");
    abr_handle h = cols_.handle();
    const bool same = (const void *)&rows_ == (const void *)&cols_;
    const double *row_pos = row_positions(same);
    KernelDesc k = k_;
    k.bind(rows_, cols_);
    const double *rpr = upload_row_radii();
    double *b = nullptr, *y = nullptr;
    detail::check_rc(h, abr_malloc(h, (void **)&b, sizeof(double) * (cols() + 1)), "evaluate");
    detail::check_rc(h, abr_malloc(h, (void **)&y, sizeof(double) * (rows() + 1)), "evaluate");
    detail::check_rc(h, abr_memcpy_h2d(h, b, rhs.data(), sizeof(double) * cols()), "evaluate");
    detail::check_rc(h, abr_memcpy_h2d(h, y, lhs.data(), sizeof(double) * rows()), "evaluate");
    detail::check_rc(h, abr_sparse_matvec(h, row_pos, rows_.size(), same ? 1 : 0, &k.d, radius_, rpr, b, y, nullptr), "evaluate");
    detail::check_rc(h, abr_memcpy_d2h(h, lhs.data(), y, sizeof(double) * rows()), "evaluate");
    abr_free(h, b);
    abr_free(h, y);
    if (rpr) abr_free(h, const_cast<double *>(rpr));
  }

  // the same on device-resident vectors: nothing is allocated or copied — the form an iterative solver
  // loop wants (two products per iteration on the same operator)
  void evaluate(DeviceVector &lhs, const DeviceVector &rhs) const {
    ABR_CHECK(lhs.size() == rows(), "lhs vector has incompatible size");
    ABR_CHECK(rhs.size() == cols(), "rhs vector has incompatible size");
    ABR_CHECK(cols_.searchable(), "column particles have no neighbour search");
    abr_handle h = cols_.handle();
    const bool same = (const void *)&rows_ == (const void *)&cols_;
    const double *row_pos = row_positions(same);
    KernelDesc k = k_;
    k.bind(rows_, cols_);
    const double *rpr = upload_row_radii();
    detail::check_rc(h, abr_sparse_matvec(h, row_pos, rows_.size(), same ? 1 : 0, &k.d, radius_, rpr, rhs.data(), lhs.data(), nullptr), "evaluate");
    if (rpr) abr_free(h, const_cast<double *>(rpr));
  }
  // y = K * b on the device
  DeviceVector operator*(const DeviceVector &b) const {
    DeviceVector y(cols_.handle(), rows());
    y.set_zero();
    evaluate(y, b);
    return y;
  }

  // y = K * b   (Eigen zeroes the destination first, src/detail/Operators.h:219-232)
  template <typename VectorType> VectorType operator*(const VectorType &b) const {
    VectorType y(rows());
    for (size_t i = 0; i < rows(); ++i) y[i] = 0.0;
    evaluate(y, b);
    return y;
  }

  // K.coeff(i, j) (src/Operators.h:149-151 -> src/Kernels.h:102-112 over detail::sparse_kernel,
  // src/detail/Kernels.h:336-367): minimum-image dx, STRICT |dx|^2 < r^2
  double coeff(const size_t i, const size_t j) const {
    ABR_CHECK(i < rows(), "i greater than rows()");
    ABR_CHECK(j < cols(), "j greater than cols()");
    abr_handle h = cols_.handle();
    const bool same = (const void *)&rows_ == (const void *)&cols_;
    const double *row_pos = row_positions(same);
    KernelDesc k = k_;
    k.bind(rows_, cols_);
    const double *rpr = upload_row_radii();
    uint64_t *ij = nullptr;
    detail::check_rc(h, abr_malloc(h, (void **)&ij, 3 * sizeof(uint64_t)), "coeff");
    const uint64_t host_ij[2] = {i, j};
    detail::check_rc(h, abr_memcpy_h2d(h, ij, host_ij, sizeof(host_ij)), "coeff");
    detail::check_rc(h, abr_query_set_particles(h, cols_.device_positions(), cols_.device_size()), "coeff");
    detail::check_rc(h, abr_sparse_coeff(h, row_pos, rows_.size(), &k.d, radius_, rpr, ij, ij + 1, 1, reinterpret_cast<double *>(ij + 2)), "coeff");
    double out = 0;
    detail::check_rc(h, abr_memcpy_d2h(h, &out, ij + 2, sizeof(out)), "coeff");
    abr_free(h, ij);
    if (rpr) abr_free(h, const_cast<double *>(rpr));
    return out;
  }

  // K.assemble(std::vector<Triplet>&, startI, startJ) (src/Kernels.h:653-685): one
  // (row, col, value) triplet per scalar entry, rows in order, the entries of a row in the
  // order of the reference's search iterator.  Triplet: any type constructible from
  // (row, col, value), e.g. Eigen::Triplet<double>.
  template <typename Triplet> void assemble(std::vector<Triplet> &triplets, const size_t startI = 0, const size_t startJ = 0) const {
    ABR_CHECK(cols_.searchable(), "column particles have no neighbour search");
    abr_handle h = cols_.handle();
    const bool same = (const void *)&rows_ == (const void *)&cols_;
    const double *row_pos = row_positions(same);
    KernelDesc k = k_;
    k.bind(rows_, cols_);
    const double *rpr = upload_row_radii();
    const size_t nr = rows_.size(), BR = k.d.block_rows, BC = k.d.block_cols;
    uint32_t *row_ptr = nullptr;
    detail::check_rc(h, abr_malloc(h, (void **)&row_ptr, (nr + 1) * sizeof(uint32_t)), "assemble");
    uint64_t nnz = 0;
    detail::check_rc(h, abr_sparse_assemble(h, row_pos, nr, same ? 1 : 0, &k.d, radius_, rpr, row_ptr, nullptr, nullptr, 0, &nnz), "assemble");
    int32_t *col = nullptr;
    double *val = nullptr;
    detail::check_rc(h, abr_malloc(h, (void **)&col, (nnz + 1) * sizeof(int32_t)), "assemble");
    detail::check_rc(h, abr_malloc(h, (void **)&val, (nnz * BR * BC + 1) * sizeof(double)), "assemble");
    detail::check_rc(h, abr_sparse_assemble(h, row_pos, nr, same ? 1 : 0, &k.d, radius_, rpr, row_ptr, col, val, nnz, &nnz), "assemble");
    std::vector<uint32_t> hp(nr + 1);
    std::vector<int32_t> hc(nnz);
    std::vector<double> hv(nnz * BR * BC);
    detail::check_rc(h, abr_memcpy_d2h(h, hp.data(), row_ptr, (nr + 1) * sizeof(uint32_t)), "assemble");
    if (nnz) {
      detail::check_rc(h, abr_memcpy_d2h(h, hc.data(), col, nnz * sizeof(int32_t)), "assemble");
      detail::check_rc(h, abr_memcpy_d2h(h, hv.data(), val, nnz * BR * BC * sizeof(double)), "assemble");
    }
    for (size_t i = 0; i < nr; ++i)
      for (uint32_t e = hp[i]; e < hp[i + 1]; ++e)
        for (size_t ii = 0; ii < BR; ++ii)
          for (size_t jj = 0; jj < BC; ++jj)
            triplets.push_back(Triplet(i * BR + ii + startI, (size_t)hc[e] * BC + jj + startJ, hv[((size_t)e * BR + ii) * BC + jj]));
    abr_free(h, row_ptr);
    abr_free(h, col);
    abr_free(h, val);
    if (rpr) abr_free(h, const_cast<double *>(rpr));
  }

  // FRadius overload (src/Operators.h:478-489): radius_function(a) per row particle, evaluated on
  // the host when the operator is applied (the reference evaluates it inside the row loop,
  // src/Kernels.h:739) and shipped as radius_per_row
  template <typename FRadius> void set_radius_function(const FRadius &f) {
    radius_fn_ = [f](const typename RowParticles::value_type &a) { return (double)f(a); };
  }

private:
  // Device positions of the row set.  Rows that are not the column set need no search structure
  // (tests/rbf_interpolation.h:326): every column of the row set is shipped as it is on the host now,
  // so that the functor's row columns (bind) are neither null nor stale.
  const double *row_positions(bool same) const {
    if (same) return cols_.device_positions();
    rows_.sync_all_to_device();
    return rows_.device_positions();
  }
  const double *upload_row_radii() const {
    if (!radius_fn_) return nullptr;
    abr_handle h = cols_.handle();
    std::vector<double> r(rows_.size() + 1);
    for (size_t i = 0; i < rows_.size(); ++i) r[i] = radius_fn_(rows_[i]);
    double *p = nullptr;
    detail::check_rc(h, abr_malloc(h, (void **)&p, sizeof(double) * r.size()), "radius");
    detail::check_rc(h, abr_memcpy_h2d(h, p, r.data(), sizeof(double) * rows_.size()), "radius");
    return p;
  }
  const RowParticles &rows_;
  const ColParticles &cols_;
  double radius_;
  KernelDesc k_;
  std::function<double(const typename RowParticles::value_type &)> radius_fn_;
};

// src/Operators.h:508-516
template <typename RowParticles, typename ColParticles, typename KernelDesc>
SparseOperator<RowParticles, ColParticles, KernelDesc> create_sparse_operator(const RowParticles &rows, const ColParticles &cols,
                                                                             const double radius, const KernelDesc &k) {
  return SparseOperator<RowParticles, ColParticles, KernelDesc>(rows, cols, radius, k);
}
// src/Operators.h:478-489: the radius as a function of the row particle
template <typename RowParticles, typename ColParticles, typename FRadius, typename KernelDesc,
          typename = typename std::enable_if<!std::is_arithmetic<FRadius>::value>::type>
SparseOperator<RowParticles, ColParticles, KernelDesc> create_sparse_operator(const RowParticles &rows, const ColParticles &cols,
                                                                             const FRadius &radius_function, const KernelDesc &k) {
  SparseOperator<RowParticles, ColParticles, KernelDesc> op(rows, cols, 0.0, k);
  op.set_radius_function(radius_function);
  return op;
}


// ---- Level 3: the neighbour sum of the symbolic layer ------------------------------------
// Symbol / Label / create_dx / AccumulateWithinDistance with the reference's names and call shape
// (src/Symbolic.h:269-444):
//     Symbol<rho> r;  Label<0, P> a(particles);  Label<1, P> b(particles);
//     AccumulateWithinDistance<std::plus<double>> sum(2 * h);
//     r[a] = sum(b, summand);            // rho_a = sum over b within 2h of a of summand(dx, a, b)
// evaluated as sparse_sum_impl does (src/detail/Contexts.h:247-289: init, then `sum = sum + expr`
// over distance_search<2>(b-set, r_a, max_distance)) — on the GPU, through the same cell-tiled product
// (b == 1, y preset to init).  The one restriction: the summand is not a Boost.Proto expression but a
// kernel function of the device (a descriptor of namespace kernels, e.g. kernels::sph_density), scalar or
// D x 1 vector valued.
template <typename Var> struct Symbol;
template <unsigned int I, typename P> struct Label {
  explicit Label(P &p) : particles(p) {}
  P &get_particles() const { return particles; }
  P &particles;
};
// dx = r_b - r_a of a pair of labels (src/Symbolic.h:372-388): a marker here — the device functors receive dx
template <typename LA, typename LB> struct Dx {};
template <unsigned int I, unsigned int J, typename P> Dx<Label<I, P>, Label<J, P>> create_dx(const Label<I, P> &, const Label<J, P> &) { return {}; }

namespace detail {
template <typename ColParticles, typename KernelDesc> struct within_distance_sum {
  const ColParticles &cols;
  double max_distance;
  KernelDesc summand;
  double init;
};
inline void store_sum(double &dst, const double *y, size_t) { dst = y[0]; }
template <unsigned int N> inline void store_sum(Vector<double, N> &dst, const double *y, size_t br) {
  for (size_t k = 0; k < br && k < N; ++k) dst[k] = y[k];
}
template <typename Var, typename P> struct symbol_ref {
  P &rows;
  // s[a] = sum(b, summand)
  template <typename ColParticles, typename KernelDesc> symbol_ref &operator=(const within_distance_sum<ColParticles, KernelDesc> &e) {
    const size_t br = e.summand.d.block_rows;
    ABR_CHECK(e.summand.d.block_cols == 1, "AccumulateWithinDistance: the summand must be scalar or D x 1");
    SparseOperator<P, ColParticles, KernelDesc> op(rows, e.cols, e.max_distance, e.summand);
    std::vector<double> y(rows.size() * br, e.init), ones(e.cols.size(), 1.0);
    op.evaluate(y, ones);
    auto &col = rows.template column<Var>();
    for (size_t i = 0; i < rows.size(); ++i) store_sum(col[i], y.data() + i * br, br);
    return *this;
  }
};
} // namespace detail

template <typename Var> struct Symbol {
  template <unsigned int I, typename P> detail::symbol_ref<Var, P> operator[](const Label<I, P> &a) const { return detail::symbol_ref<Var, P>{a.get_particles()}; }
};

// AccumulateWithinDistance<std::plus<T>> (src/Symbolic.h:420-444); other functors are not offered on the device
template <typename T> class AccumulateWithinDistance;
template <typename T> class AccumulateWithinDistance<std::plus<T>> {
public:
  explicit AccumulateWithinDistance(const double max_distance = 1.0) : max_distance_(max_distance), init_(0.0) {}
  void set_max_distance(const double max_distance) { max_distance_ = max_distance; }
  void set_init(const double init) { init_ = init; }
  template <unsigned int I, typename P, typename KernelDesc> detail::within_distance_sum<P, KernelDesc> operator()(const Label<I, P> &b, const KernelDesc &summand) const {
    return detail::within_distance_sum<P, KernelDesc>{b.get_particles(), max_distance_, summand, init_};
  }

private:
  double max_distance_, init_;
};

// ---- zero and block operators -------------------------------------------------------
// create_zero_operator (src/Operators.h:531-537, KernelZero)
template <typename RowParticles, typename ColParticles> class ZeroOperator {
public:
  ZeroOperator(const RowParticles &rows, const ColParticles &cols) : rows_(rows), cols_(cols) {}
  size_t rows() const { return rows_.size(); }
  size_t cols() const { return cols_.size(); }
  template <typename LHS, typename RHS> void evaluate(LHS &, const RHS &) const {}
  double coeff(size_t, size_t) const { return 0.0; }
  template <typename Triplet> void assemble(std::vector<Triplet> &, size_t = 0, size_t = 0) const {}

private:
  const RowParticles &rows_;
  const ColParticles &cols_;
};
template <typename RowParticles, typename ColParticles>
ZeroOperator<RowParticles, ColParticles> create_zero_operator(const RowParticles &rows, const ColParticles &cols) {
  return ZeroOperator<RowParticles, ColParticles>(rows, cols);
}

// create_block_operator<NI,NJ>(blocks...) (src/Operators.h:541-548): NI x NJ blocks, row
// major, behind one operator.  Product as in src/detail/Operators.h:170-198 (block (I,J)
// works on y.segment(start_row(I)) and x.segment(start_col(J)) and accumulates), coeff as
// in src/Operators.h:242-251, assemble with the block's start offsets (:253-262).
namespace detail {
// a window of a vector with data()/size()/operator[]
struct segment {
  double *p;
  size_t n;
  double *data() { return p; }
  const double *data() const { return p; }
  size_t size() const { return n; }
  double &operator[](size_t i) { return p[i]; }
  const double &operator[](size_t i) const { return p[i]; }
};
} // namespace detail

template <unsigned int NI, unsigned int NJ, typename... Ops> class BlockOperator {
  static_assert(sizeof...(Ops) == NI * NJ, "create_block_operator: need NI*NJ blocks");

public:
  explicit BlockOperator(const Ops &... ops) : blocks_(ops...) {}
  size_t rows() const { return row_start(NI); }
  size_t cols() const { return col_start(NJ); }

  template <typename LHS, typename RHS> void evaluate(LHS &lhs, const RHS &rhs) const {
    ABR_CHECK((size_t)lhs.size() == rows(), "lhs vector has incompatible size");
    ABR_CHECK((size_t)rhs.size() == cols(), "rhs vector has incompatible size");
    eval_all(lhs, rhs, std::make_index_sequence<NI * NJ>());
  }
  template <typename VectorType> VectorType operator*(const VectorType &b) const {
    VectorType y(rows());
    for (size_t i = 0; i < rows(); ++i) y[i] = 0.0;
    evaluate(y, b);
    return y;
  }
  double coeff(const size_t i, const size_t j) const {
    double out = 0.0;
    coeff_all(i, j, out, std::make_index_sequence<NI * NJ>());
    return out;
  }
  template <typename Triplet> void assemble(std::vector<Triplet> &triplets) const {
    assemble_all(triplets, std::make_index_sequence<NI * NJ>());
  }

private:
  template <size_t K> size_t block_rows() const { return std::get<K>(blocks_).rows(); }
  template <size_t K> size_t block_cols() const { return std::get<K>(blocks_).cols(); }
  // sizes of block row I = rows of block (I, 0); of block column J = cols of block (0, J)
  size_t rows_of(size_t I) const { return rows_of_impl(I, std::make_index_sequence<NI * NJ>()); }
  size_t cols_of(size_t J) const { return cols_of_impl(J, std::make_index_sequence<NI * NJ>()); }
  template <size_t... K> size_t rows_of_impl(size_t I, std::index_sequence<K...>) const {
    size_t r = 0;
    int dummy[] = {((K == I * NJ) ? (r = block_rows<K>(), 0) : 0)...};
    (void)dummy;
    return r;
  }
  template <size_t... K> size_t cols_of_impl(size_t J, std::index_sequence<K...>) const {
    size_t c = 0;
    int dummy[] = {((K == J) ? (c = block_cols<K>(), 0) : 0)...};
    (void)dummy;
    return c;
  }
  size_t row_start(size_t I) const {
    size_t s = 0;
    for (size_t i = 0; i < I; ++i) s += rows_of(i);
    return s;
  }
  size_t col_start(size_t J) const {
    size_t s = 0;
    for (size_t j = 0; j < J; ++j) s += cols_of(j);
    return s;
  }
  template <typename LHS, typename RHS, size_t... K> void eval_all(LHS &lhs, const RHS &rhs, std::index_sequence<K...>) const {
    int dummy[] = {(eval_one<K>(lhs, rhs), 0)...};
    (void)dummy;
  }
  template <size_t K, typename LHS, typename RHS> void eval_one(LHS &lhs, const RHS &rhs) const {
    const size_t I = K / NJ, J = K % NJ;
    detail::segment y{lhs.data() + row_start(I), rows_of(I)};
    const detail::segment x{const_cast<double *>(rhs.data()) + col_start(J), cols_of(J)};
    std::get<K>(blocks_).evaluate(y, x);
  }
  template <size_t... K> void coeff_all(size_t i, size_t j, double &out, std::index_sequence<K...>) const {
    int dummy[] = {(coeff_one<K>(i, j, out), 0)...};
    (void)dummy;
  }
  template <size_t K> void coeff_one(size_t i, size_t j, double &out) const {
    const size_t I = K / NJ, J = K % NJ;
    if (i >= row_start(I) && i < row_start(I + 1) && j >= col_start(J) && j < col_start(J + 1))
      out += std::get<K>(blocks_).coeff(i - row_start(I), j - col_start(J));
  }
  template <typename Triplet, size_t... K> void assemble_all(std::vector<Triplet> &t, std::index_sequence<K...>) const {
    int dummy[] = {(std::get<K>(blocks_).assemble(t, row_start(K / NJ), col_start(K % NJ)), 0)...};
    (void)dummy;
  }
  std::tuple<Ops...> blocks_;
};
template <unsigned int NI, unsigned int NJ, typename... Ops> BlockOperator<NI, NJ, Ops...> create_block_operator(const Ops &... ops) {
  return BlockOperator<NI, NJ, Ops...>(ops...);
}

} // namespace Aboria
#endif
