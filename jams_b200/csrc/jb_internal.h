// jb_internal.h — context, device-side parameter blocks and launcher prototypes shared by
// jb_capi.cu (host side of the C ABI) and jb_kernels.cu (sm_100a kernels).
//
// Device data layout (DESIGN.md "Data layout in HBM"):
//   spins are stored SoA in double, one array per component, in a GHOSTED box
//       index(xp, yp, m, zp) = ((xp * PY + yp) * M + m) * PZ + zp
//   with xp = x + gx, yp = y + gy, zp = z + oz and ghost depths gx,gy,gz = max |T| of the exchange
//   template along each axis.  oz (4 by default; 8 / 16 selectable) >= gz keeps every
//   interior run on a 32-byte sector boundary (all stores write whole sectors) with the shortest possible gap between rows, and
//   TMA boxes start on even columns (a box whose first element is not 16-byte aligned faults).  z is the fastest index (lanes of a warp run along z), the motif index
//   m sits between y and z so that a warp never mixes motif sites.  Ghost cells hold the periodic
//   image (or zero across an open boundary: a zero spin contributes nothing to J.s), so the field
//   gather has no boundary branches and a tile + halo is a plain box for TMA.
//   The reference orders sites ((x*Ny + y)*Nz + z)*M + m (core/lattice.cc:622-657) in AoS N x 3;
//   jb_import_spins / jb_export_spins translate.
#ifndef JB_INTERNAL_H
#define JB_INTERNAL_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "jams_b200.h"

#define JB_MAX_MOTIF 32
#define JB_MAX_CLASSES 255
#define JB_MAX_RING 8
#define JB_COPY_CHUNKS 8   // x-chunks of the pipelined host <-> device spin transfer

// ---- geometry of the ghosted box ---------------------------------------------------------------
struct JbGeom {
  int nx, Ny, Nz, M;       // interior extent of this slab (cells) and motif size
  int gx, gy, gz;          // ghost depth
  int oz;                  // column of z = 0 inside a row: 4, 8 or 16 doubles (interior rows start on a sector / DRAM atom / L2 line)
  int PX, PY, PZ;          // padded extent (PZ a multiple of oz)
  long long sY;            // stride of yp  = M * PZ
  long long sX;            // stride of xp  = PY * M * PZ
  long long elems;         // PX * sX
  int per[3];              // periodic flags
  int x_begin;             // global x of local x = 0
  int Nx_global;
  int n_ranks, rank;
};

// ---- per-class constants (a class = one distinct tuple of per-site parameters) -----------------
// The reference keeps all of these as per-site arrays (globals::mus/gyro/alpha, uniaxial magnitude_/axis_,
// zeeman dc_local_field_); they only ever take one value per material / motif position, so they are
// de-duplicated on the host and cost 0 B of HBM traffic per spin.
struct JbClass {
  double inv_mu;      // 1 / mu_i
  double mu;          // mu_i (for applied field and magnetisation)
  double c_full;      // -gyro_i * dt      (set per jb_step)
  double c_half;      // -gyro_i * dt / 2
  double alpha;
  double sigma;       // sqrt(2 kB alpha / (mu gyro dt [1+alpha^2])) * sqrt(T)   (set per jb_step)
  double Kp;          // K_i * power, meV
  double K;           // K_i, meV
  double KpT;         // K_i * power / mu_i, Tesla
  double ax, ay, az;  // anisotropy axis
  double fx, fy, fz;  // constant field of this stage, meV: dc + ac*cos(omega t) + mu*B_applied
  double fTx, fTy, fTz;  // the same divided by mu_i, Tesla
  int power;          // 0 = no uniaxial term
  int pad;
  double gyro;        // gyro_i (rad / ps / T): the RK4 stages need k = -gyro (...) without the step folded in
};

// A second / third uniaxial Hamiltonian of a configuration (the reference sums any number, e.g. K1 and K2 as two modules).  Kept out
// of JbClass on purpose: the TMA stage kernels take their class constants as kernel parameters and never see these; a context
// that has any runs its steps on the direct-gather kernels (choose_tiling), which read this table by class id.
#define JB_MAX_UNIAXIAL 3
struct JbUniExtra {
  double Kp;          // K_i * power, meV
  double K;           // K_i, meV
  double KpT;         // K_i * power / mu_i, Tesla
  double ax, ay, az;  // anisotropy axis
  int power;          // 0 = this class has no such term
  int pad;
};

// one entry of the exchange template of a motif site, in ghosted-box terms
struct JbNbr {
  int delta;   // offset inside a plane of the ghosted box: (dy*M + (mj - mi))*PZ + dz
  int dx;      // plane offset
  int jidx;    // index into the table of unique tensors
  int pad;
  double J;    // scalar coupling (isotropic case), meV
};

struct JbTables {
  const JbNbr *nbr_global;     // entries with delta computed for the global ghosted box (row = PZ)
  const double *Jtab;          // n_unique x 9
  const JbClass *classes;      // n_classes
  const unsigned char *site_class;  // per-site class in interior [x][y][m][z] order, or nullptr (motif-uniform)
  int nbr_begin[JB_MAX_MOTIF + 1];
  int class_of_motif[JB_MAX_MOTIF];
  int n_classes;
  int iso;                     // all tensors are scalar multiples of the identity
  // biquadratic exchange (hamiltonian/cuda_biquadratic_exchange_kernel.cuh): its own template, J = B_ij in meV; null = none
  const JbNbr *bq_global;
  int bq_begin[JB_MAX_MOTIF + 1];
  // uniaxial slots 1 and 2 (jb_set_uniaxial_term): [class * (JB_MAX_UNIAXIAL - 1) + slot - 1]; null = none
  const JbUniExtra *uni_extra;
};

// ---- parameter block of the fused stage kernels --------------------------------------------------
struct JbStageParams {
  JbGeom g;
  JbTables t;
  const double *in[3];   // spins read with neighbours (S0 in stage A, S1 in stage B)
  double *out[3];        // spins written (S1 in stage A, S0 in stage B), own box
  double *out_lo[3];     // box that receives the images of my low-x boundary planes (own box, a peer's, or null)
  double *out_hi[3];     // ... of my high-x boundary planes
  double *u[3];          // Heun intermediate u = s_n + dt/2 k1 (written in A, read in B), interior only used
  unsigned long long seed, step;
  int thermal;
  // RK4 (jb_step_rk4): the state at the start of the step (own site only), the step, and where the running sum
  // k1 + 2 k2 + 2 k3 lives is `u`
  const double *s_old[3];
  double dt;
};

// ---- parameter block of the persistent TMA stage kernel (jb_stage_pair.cu) ------------------------------
// The per-class constants, the template's group offsets, the x-chunk plan and the tiling live in the kernel parameter
// (constant) bank; the exchange template itself (16 B per entry) is copied from global to shared memory once per
// resident CTA.  "Matrix" data therefore costs 0 B of HBM traffic per spin (the reference streams 12 B per non-zero,
// containers/sparse_matrix.h:366-379).
#define JB_TILE_MAX_NBR 1024
#define JB_TILE_MAX_CLASSES 8
#define JB_TILE_MAX_MOTIF 16
#define JB_TILE_MAX_GX 3
#define JB_PAIR_MAX_RING 12   // ring depth limit of the stage kernel
#define JB_TILE_MAX_CHUNKS 160// x-chunks of the work-item plan
#define JB_TRACE_WORDS 32     // per-CTA trace record: {SM, first ns, last ns, items, then (item << 40 | start ns - first) per item}
#define JB_TRACE_ITEMS (JB_TRACE_WORDS - 4)
#define JB_ITEM_RING 16       // item ids in flight between the producer warp and the consumer warps (> JB_PAIR_MAX_RING)
struct __align__(16) JbTileNbr {
  int delta;   // offset inside a plane of the smem tile: (dy*M + (mj - mi))*BZ + dz
  int d;       // dx + gx: which of the 2 gx + 1 resident planes
  double J;    // scalar coupling divided by mu of the owning motif site: Tesla
};
// halo handshake of a slab-decomposed run, done INSIDE the stage kernel (no separate wait / signal launches): the producer
// thread of an item that loads ghost planes polls this rank's flag for the neighbour's previous stage, and the last consumer
// warp that finishes the face items of a side publishes this stage's epoch in the neighbour's flag (jb_stage_pair.cu)
struct JbHalo {
  unsigned long long *flags;      // this rank's flags: [0] written by the lo neighbour, [1] by the hi neighbour, [2] error
  unsigned long long *sig_lo;     // the lo neighbour's flag that I write (its [1]), or null
  unsigned long long *sig_hi;     // the hi neighbour's flag that I write (its [0]), or null
  unsigned long long wait_epoch;  // both neighbours must have published this epoch before I touch ghost planes
  unsigned long long signal_epoch;
  unsigned int *face_count;       // [0] lo, [1] hi: finished face items, counted once per CTA and item (reset by the last one)
  unsigned int face_target[2];    // face items per side
  int enabled;
};
// one segment of the exchange template as the rows kernel (jb_stage_rows.cu) sees it: up to five entries of a motif site that
// differ only in their y offset, dy = dy0 ... dy0 + L - 1 (absent offsets carry a zero coupling)
struct __align__(16) JbRowSeg {
  int delta;    // byte offset of the neighbour of the thread's FIRST site at dy0 (component x) relative to that site, inside a slot
  int d;        // dx + gx: which of the 2 gx + 1 resident planes
  int L;        // 1 ... 5
  int pad;
  double c[6];  // couplings / mu_i (Tesla) of dy0 + t, t < L; the rest zero
};
#define JB_ROWS_MAX_L 5
#define JB_ROWS_Q 4           // sites (consecutive y rows) per consumer thread
#define JB_ROWS_MAX_WARPS 8   // consumer warps per CTA (+ the producer warp: 288 threads, 168 registers per thread)
struct JbTileParams {
  JbGeom g;
  double *out[3];        // spins written (S1 in stage A, S0 in stage B), own box
  double *out_lo[3];     // box that receives the images of my low-x boundary planes (own box, a peer's, or null)
  double *out_hi[3];
  double *u[3];          // Heun intermediate, written by stage A with plain stores (stage B reads it through TMA)
  const double *J9T;     // n_nbr x 9: tensor of every template entry divided by mu_i, Tesla (anisotropic exchange only)
  unsigned long long step;
  double dt;             // RK4 stages: the step (the Heun stages have it folded into the class constants)
  unsigned int rk[20];   // Philox4x32-10 round keys of the seed (k0 + r W0, k1 + r W1)
  int TY, TZ, UZ;        // tile extent in y, z; UZ = TZ rounded up to even = inner extent of the u box
  int BY, BZ, gzb;       // tile + halo extent; gzb = gz rounded up to even = z halo of the box (BZ = TZ + 2 gzb)
  int slotS, slotU;      // doubles per component per ring slot (multiples of 16 = 128 B)
  int R, RU;             // ring depths: S planes (>= 2 gx + 2), U planes (>= 2)
  int msplit;            // consumer threads per z pair of a y row: thread (ty, ms, zp) owns the motif sites ms, ms + msplit, ... (a divisor of M)
  int nbr_odd[JB_TILE_MAX_MOTIF];    // [m] -> first entry of nbr[] with an odd z offset (entries are even-first per motif site)
  // isotropic templates: the couplings (Tesla) of the entries (0, 0, -1) and (0, 0, +1) to the same motif site are taken out of
  // the table -- one of the two neighbours of each site of a pair is the other site of the pair, already in registers
  double zself[JB_TILE_MAX_MOTIF][2];
  int has_zself[JB_TILE_MAX_MOTIF];
  int n_yt, n_zt, n_cols, n_chunks, n_items;
  // work queue: items are handed out by an atomic counter in the order of this plan; item = chunk * n_cols + column.  The plan
  // lists the slab's two face chunks first (their halo traffic and flags go out early), then long chunks, then short ones (tail)
  int chunk_x0[JB_TILE_MAX_CHUNKS], chunk_xc[JB_TILE_MAX_CHUNKS];
  unsigned int *queue;        // next item of this launch
  unsigned int *queue_next;   // the counter the next launch on the stream will use: zeroed by this one
  JbHalo halo;
  unsigned long long *trace;   // optional: per CTA JB_TRACE_WORDS words (option "trace")
  // rows kernel: the segments in global memory (copied to shared memory once per CTA), sorted by motif site and length:
  // segments of motif site m with L entries are row_begin[m][L - 1] ... row_begin[m][L] - 1
  const JbRowSeg *rows;
  int n_rows;
  int row_begin[JB_TILE_MAX_MOTIF][JB_ROWS_MAX_L + 1];
  int n_nbr;
  const JbTileNbr *nbr;  // n_nbr entries in global memory, copied to shared memory once per CTA
  int nbr_begin[JB_TILE_MAX_MOTIF + 1];  // [m] -> first entry of nbr[]
  JbClass cls[JB_TILE_MAX_MOTIF];   // constants of motif site m (already resolved through class_of_motif): for M == 1 a
                                     // compile-time offset, i.e. constant-bank operands of the fp64 instructions
};

struct jb_ctx {
  jb_lattice_desc d{};
  JbGeom g{};
  int N = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;

  // host copies of the caller's per-site parameters (kept to build classes lazily)
  std::vector<double> h_mus, h_gyro, h_alpha;
  std::vector<double> h_K, h_axis; int uni_power = 0;
  std::vector<double> h_Kx[JB_MAX_UNIAXIAL - 1], h_axisx[JB_MAX_UNIAXIAL - 1]; int uni_powerx[JB_MAX_UNIAXIAL - 1] = {0, 0};   // slots 1, 2
  std::vector<JbUniExtra> h_uni_extra; JbUniExtra *d_uni_extra = nullptr;
  bool has_uni_extra() const { return uni_powerx[0] != 0 || uni_powerx[1] != 0; }
  std::vector<double> h_dc, h_ac, h_omega; bool has_zeeman = false, has_ac = false;
  double applied_B[3] = {0, 0, 0}; bool has_applied = false;
  int applied_type = 0; double applied_t0 = 0.0, applied_fbw = 0.0, applied_fc = 0.0;   // jb_set_applied_field_pulse

  // exchange template (host)
  std::vector<int> t_mi, t_mj, t_T; std::vector<double> t_J9;
  bool has_template = false;
  // biquadratic exchange template (host): scalar B per entry
  std::vector<int> bq_mi, bq_mj, bq_T; std::vector<double> bq_B;
  bool has_bq = false, bq_built = false;
  JbNbr *d_bq_global = nullptr;
  int bq_begin[JB_MAX_MOTIF + 1] = {0};
  // exchange pairs (host -> device ELL), general path
  bool has_pairs = false;
  int pairs_reach_x = 0;          // several ranks: largest x distance of a pair = depth of the x ghost planes the list addresses
  int ell_width = 0;
  int *d_ell_idx = nullptr;       // width x N (column-major: entry e of site q at e*N + q), ghosted index or -1
  int *d_ell_val = nullptr;       // value ids
  double *d_pair_J = nullptr;     // n_values x 9
  int n_pair_values = 0;
  bool pairs_iso = false;

  // classes
  bool classes_dirty = true;
  std::vector<double> class_sig;            // what the device class table currently encodes
  std::vector<JbClass> h_classes;          // without sigma / f (filled per launch)
  std::vector<double> h_class_gyro;        // gyro of every class
  std::vector<double> h_class_dc, h_class_ac, h_class_omega;  // per class x3 / x3 / x1
  std::vector<unsigned char> h_site_class; // interior [x][y][m][z] order
  bool motif_uniform = true;
  int class_of_motif[JB_MAX_MOTIF] = {0};

  // device tables
  JbNbr *d_nbr_global = nullptr;
  double *d_Jtab = nullptr;
  JbClass *d_classes = nullptr;
  unsigned char *d_site_class = nullptr;
  int nbr_begin[JB_MAX_MOTIF + 1] = {0};
  int n_unique_J = 0;
  bool iso = true;
  bool tables_built = false;

  // tiling of the persistent TMA stage kernel (jb_capi.cu choose_tiling) and its parameter-bank tables
  struct Tiling {
    bool ok = false;
    bool rows = false;                    // the rows kernel (jb_stage_rows.cu) instead of the pair kernel
    int TY = 0, TZ = 0, UZ = 0, BY = 0, BZ = 0, gzb = 0, slotS = 0, slotU = 0, R = 0, RU = 0;
    int Rs[2] = {0, 0};                   // ring depth per stage
    int n_yt = 0, n_zt = 0, n_cols = 0, threads = 0, msplit = 1;
    size_t smem[2] = {0, 0};              // per stage
    // launch shape per kernel variant [stage][thermal][recover_u]: resident CTAs and the x-chunk plan (0 = not determined yet)
    struct Shape { int grid = 0, n_chunks = 0, face_items[2] = {0, 0}; int x0[JB_TILE_MAX_CHUNKS], xc[JB_TILE_MAX_CHUNKS]; };
    Shape shape[2][2][2];
    Shape shape_rk4[4][2];                // the RK4 stages on the pair kernel [stage][thermal]
  } tiling;
  bool tiling_valid = false;
  std::vector<int> tile_order, tile_jidx;   // template entries in table order / their unique-tensor ids
  std::vector<JbClass> h_class_tab;         // host copy of the class table(s) last uploaded (parameter bank of the tile kernel)
  std::vector<JbTileNbr> tile_nbr;
  JbTileNbr *d_tile_nbr = nullptr;
  double *d_tile_J9T = nullptr;
  std::vector<int> tile_nbr_begin, tile_nbr_odd;
  std::vector<JbRowSeg> row_segs;           // rows kernel: the template cut into y segments (build_row_segments)
  JbRowSeg *d_rows = nullptr;
  int row_begin[JB_TILE_MAX_MOTIF][JB_ROWS_MAX_L + 1] = {{0}};
  std::vector<double> tile_zself;   // [m][2], see JbTileParams::zself
  int num_sms = 0;

  // state: ghosted SoA arrays
  double *S0[3] = {nullptr, nullptr, nullptr};
  double *S1[3] = {nullptr, nullptr, nullptr};
  double *U[3] = {nullptr, nullptr, nullptr};
  double *V[3] = {nullptr, nullptr, nullptr};     // second stage-input box of the RK4 solver (in the slab, so neighbours can store into it)
  bool state_allocated = false;
  double *d_aos = nullptr; size_t d_aos_bytes = 0;     // staging for host <-> device AoS
  double *d_scratch = nullptr; size_t d_scratch_bytes = 0;  // per-spin scalars / reductions
  double *h_pinned = nullptr; size_t h_pinned_bytes = 0;
  cudaStream_t copy_stream = nullptr;      // host <-> device copies of jb_import_spins / jb_export_spins, overlapped with the layout kernels
  cudaEvent_t copy_ev[JB_COPY_CHUNKS + 1] = {nullptr};

  // TMA descriptors: [0] = S0 x,y,z  [1] = S1 x,y,z (tile + halo boxes)  [2] = U x,y,z (tile boxes)
  // [3] = S0, [4] = S1 with the tile box of U (recover_u: the corrector reads the site's own s_n); [5] = V, tile + halo boxes (RK4)
  CUtensorMap tmap[6][3];
  bool tmap_valid = false;

  // options
  int opt_kernel = 2;      // 0 = direct global gathers, 2 = TMA pair kernel where the template allows it (default)
  int reach[3] = {0, 0, 0};   // max |T| of the exchange template per axis
  int opt_TY = 0, opt_TZ = 0, opt_R = 0, opt_RU = 0, opt_ctas_per_sm = 0, opt_msplit = 0;  // 0 = heuristic
  int opt_chunks = 0;         // x-chunk plan: 0 = heuristic (long chunks + a tail of short ones), n > 0 = n equal chunks
  int opt_face_after = -1;    // queue the slab's face chunks after this many interior chunks (0 = first; -1 = 2 on slab-decomposed runs, else 0)
  int opt_chunk_long = 0, opt_chunk_short = 0, opt_tail_pct = -1;   // heuristic overrides: planes per long / short chunk, share of the planes in short chunks
  int opt_verbose = 0;
  int opt_grid = 0;           // upper limit of the number of resident CTAs of the persistent kernel (0 = occupancy x SMs)
  int opt_rows_warps = 0;     // rows kernel: upper limit of consumer warps per CTA (0 = JB_ROWS_MAX_WARPS)
  int opt_recover_u = 1;      // pair kernel: 1 = no stored Heun intermediate (120 B per update), 0 = store u (144 B)
  int opt_check_symmetry = 1; // refuse an exchange matrix that is not symmetric, like the reference (settings key check_sparse_matrix_symmetry)
  int opt_trace = 0;          // per-CTA {SM, first clock, last clock, items} of the last stage launch (jb_last_trace)
  int opt_fold_halo = 1;      // slab-decomposed runs: epoch handshake inside the stage kernel (1; 2 = even when a neighbour shares this GPU) or as separate wait / signal launches (0)
  int opt_oz = 4;             // column of z = 0 inside a row (4, 8 or 16 doubles): 4 = 32-byte sectors, the shortest gap between rows
  bool state_relayout = false; // an option that changes the box layout was set: re-layout at the next ensure_ready
  int opt_detect_template = 1;   // jb_set_exchange_pairs: turn translation-invariant lists into a template
  int opt_time_kernels = 0;      // N > 0: bracket the stage launches of every N-th step of a jb_step / jb_step_rk4 call with events
  bool time_this_step = false; long long timed_steps = 0;

  // halo peers
  double *peer_lo_S0[3] = {nullptr, nullptr, nullptr}, *peer_lo_S1[3] = {nullptr, nullptr, nullptr};
  double *peer_hi_S0[3] = {nullptr, nullptr, nullptr}, *peer_hi_S1[3] = {nullptr, nullptr, nullptr};
  double *peer_lo_V[3] = {nullptr, nullptr, nullptr}, *peer_hi_V[3] = {nullptr, nullptr, nullptr};
  unsigned long long *flags = nullptr;           // [0] = written by lo neighbour, [1] = by hi neighbour, [2] = error
  unsigned long long *peer_lo_flags = nullptr, *peer_hi_flags = nullptr;
  void *peer_lo_base = nullptr, *peer_hi_base = nullptr;  // mapped IPC allocations
  bool same_peer = false;
  void *slab = nullptr; size_t slab_bytes = 0;   // single allocation holding S0,S1 and flags (one IPC handle)
  unsigned long long epoch = 0;
  bool halo_connected = false;
  bool peer_on_my_device = false;   // a neighbour slab of this process lives on the same GPU (tests)

  int *d_groups = nullptr; int n_groups_set = 0;   // jb_set_magnetisation_groups

  // regions of spins (jb_set_region): device copies of the site lists
  int *d_region[JB_MAX_REGIONS] = {nullptr}; int region_n[JB_MAX_REGIONS] = {0};

  // work queue and face counters of the stage kernel: {queue[2], face_count[2]} unsigned ints, zeroed once
  unsigned int *d_queue = nullptr;
  unsigned long long stage_launches = 0;   // parity selects which of the two queue counters a launch uses
  unsigned long long *d_trace = nullptr; int trace_ctas = 0;

  // bookkeeping
  long long launches = 0;
  int last_stage_kernel = -1;   // jb_stage_kernel
  std::vector<cudaEvent_t> ev; size_t ev_used = 0;
  std::vector<int> ev_kind;
};

// ---- launchers implemented in jb_kernels.cu (all enqueue on `stream`) -----------------------------
cudaError_t jbk_import(const JbGeom &g, const double *aos, double *const dst[3], bool fill_x_ghosts, cudaStream_t stream);
cudaError_t jbk_export(const JbGeom &g, const double *const src[3], double *aos, cudaStream_t stream);
// the same for the interior planes [x_begin, x_end) only (the host path pipelines x-chunks with the PCIe copies); jbk_import =
// all planes + jbk_fill_ghosts (ghost cells <- the interior cells they are periodic images of)
cudaError_t jbk_import_planes(const JbGeom &g, const double *aos, double *const dst[3], int x_begin, int x_end, cudaStream_t stream);
cudaError_t jbk_fill_ghosts(const JbGeom &g, double *const dst[3], bool fill_x_ghosts, cudaStream_t stream);
cudaError_t jbk_export_planes(const JbGeom &g, const double *const src[3], double *aos, int x_begin, int x_end, cudaStream_t stream);
cudaError_t jbk_push_x_ghosts(const JbGeom &g, const double *const src[3], double *const lo[3], double *const hi[3], cudaStream_t stream);
cudaError_t jbk_stage_direct(const JbStageParams &p, int stage, cudaStream_t stream);
// one of the four RK4 stages (stage 0..3), direct gathers: in = stage input (with neighbours), out = next stage input
// (stages 0-2) or the new spins (stage 3), u = running sum of the k's, s_old = spins at the start of the step
cudaError_t jbk_rk4_stage_direct(const JbStageParams &p, int stage, cudaStream_t stream);
// persistent TMA stage kernel (jb_stage_pair.cu): tmaps = {S.x, S.y, S.z, U.x, U.y, U.z}; grid = number of resident CTAs;
// `threads` = consumer threads = ceil(TZ/2) x TY (the launch adds the producer warp); recu = recover_u data flow
cudaError_t jbk_stage_pair(const JbTileParams &p, const CUtensorMap *tmaps6, int stage, int thermal, int iso, int recu,
                           int threads, int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t jbk_stage_pair_occupancy(const JbTileParams &p, int stage, int thermal, int iso, int recu, int threads,
                                     size_t smem_bytes, int *blocks_per_sm);
// the four RK4 stages on the same kernel (isotropic couplings): tmaps6 = {stage input x,y,z, S0 (s_old) tile boxes x,y,z}
cudaError_t jbk_rk4_stage_pair(const JbTileParams &p, const CUtensorMap *tmaps6, int stage, int thermal, int threads, int grid,
                               size_t smem_bytes, cudaStream_t stream);
cudaError_t jbk_rk4_stage_pair_occupancy(const JbTileParams &p, int stage, int thermal, int threads, size_t smem_bytes, int *blocks_per_sm);
// rows kernel (jb_stage_rows.cu): tmaps = {S.x, S.y, S.z}; threads = 32 x (TY / 4) x msplit consumer threads
cudaError_t jbk_stage_rows(const JbTileParams &p, const CUtensorMap *tmaps3, int stage, int thermal, int threads, int grid,
                           size_t smem_bytes, cudaStream_t stream);
cudaError_t jbk_stage_rows_occupancy(int stage, int thermal, int threads, size_t smem_bytes, int *blocks_per_sm);
cudaError_t jbk_stage_pairs(const JbStageParams &p, const int *ell_idx, const int *ell_val, int width, const double *pairJ,
                            int iso, int stage, cudaStream_t stream);
// term field (meV) into AoS N x 3 (device); term as jb_term; pairs path when ell_idx != nullptr
cudaError_t jbk_field(const JbGeom &g, const JbTables &t, const double *const s[3], int term, const int *ell_idx,
                      const int *ell_val, int width, const double *pairJ, int pairs_iso, double *h_aos, cudaStream_t stream);
// per-spin energies (N, reference site order) and block-reduced total into total_out[0]; scratch >= 1024 doubles
cudaError_t jbk_energy(const JbGeom &g, const JbTables &t, const double *const s[3], int term, const int *ell_idx,
                       const int *ell_val, int width, const double *pairJ, int pairs_iso, double *e_out, double *scratch,
                       double *total_out, cudaStream_t stream);
// sum mu_i s_i and sum mu_i per group -> out4 (n_groups x 4, device); scratch >= 4*n_groups*1024 doubles
cudaError_t jbk_magnetisation(const JbGeom &g, const JbTables &t, const double *const s[3], int n_groups,
                              const int *group_of_spin, double *scratch, double *out4, cudaStream_t stream);
// sum mu_i s_i and sum mu_i over the sites of a region -> out4 (device); scratch >= 4 * 1024 doubles
cudaError_t jbk_region_moment(const JbGeom &g, const JbTables &t, const double *const s[3], const int *sites, int n,
                              double *scratch, double *out4, cudaStream_t stream);
// s_i <- R s_i for the sites of a region, ghost images included (lo / hi: boxes that receive the x images, or null)
cudaError_t jbk_region_rotate(const JbGeom &g, double *const s[3], double *const lo[3], double *const hi[3], const int *sites, int n,
                              const double R9[9], cudaStream_t stream);
cudaError_t jbk_noise(const JbGeom &g, const JbTables &t, unsigned long long seed, unsigned long long step,
                      int normals_only, double *xi_aos, cudaStream_t stream);
cudaError_t jbk_signal(unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag, unsigned long long epoch, cudaStream_t stream);
cudaError_t jbk_wait(unsigned long long *flags, int wait_lo, int wait_hi, unsigned long long epoch, cudaStream_t stream);

#endif  // JB_INTERNAL_H
