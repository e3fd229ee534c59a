/* mps_b200.h -- C ABI of the B200-native matrix-product-state gate engine (libmps_b200.so).
 *
 * This is the drop-in boundary for the tensor back end of TNQVM's `exatn-mps` visitor
 * (reference: tnqvm/visitors/exatn-mps/ExaTnMpsVisitor.{hpp,cpp}).  Every entry point below
 * names the reference call site it replaces.  Plain pointers and sizes only; complex128 is an
 * interleaved double[2]; all tensors are column-major, site index order (left bond, physical,
 * right bond) exactly as the reference keeps them inside ExaTN (ExaTnMpsVisitor.cpp:1684-1695).
 *
 * Conventions (SURVEY.md section 8a cheat-sheet, verified against the reference's gtests):
 *   - qubit 0 is the least-significant bit of a state-vector index (ExaTnMpsVisitor.cpp:1674-1682)
 *   - 1q gate: new[b] = sum_i m[b][i] old[i], m row-major as in tnqvm/base/Gates.hpp
 *   - 2q gate: m row-major 4x4, index = 2*bit(q0)+bit(q1) whichever of q0,q1 is the left site
 *     (ExaTnMpsVisitor.cpp:1492-1499); |q0-q1| must be 1 (assert at :1398)
 *   - "exp-val-z" and sampling use the raw, un-normalised state (ExaTnMpsVisitor.cpp:616-644)
 *
 * Error model: every function returns 0 on success, non-zero on failure; mps_last_error() gives
 * the message.  No C++ exception crosses this boundary.  A handle is confined to one host thread;
 * work is asynchronous on the handle's CUDA stream until a read-back or mps_sync().
 * There is no CPU fallback: mps_create fails when no CUDA device is usable.
 */
#ifndef MPS_B200_H_
#define MPS_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mps_b200_handle* mps_handle_t;

/* gauge of the SVD write-back */
#define MPS_GAUGE_REFERENCE 0 /* Q_lo = U sqrt(S), Q_hi = sqrt(S) V^H : ExaTN SVDLR, ExaTnMpsVisitor.cpp:1623 */
#define MPS_GAUGE_LEFT 1      /* Q_lo = U,        Q_hi = S V^H  (orthogonality centre moves right)           */
#define MPS_GAUGE_RIGHT 2     /* Q_lo = U S,      Q_hi = V^H    (orthogonality centre moves left)            */

/* ExatnMpsVisitor::initialize (ExaTnMpsVisitor.cpp:173-346): |0...0> with all bonds 1.
 * max_bond <= 0 -> no limit ("max-bond-dim", :265-271); svd_cutoff < 0 -> DBL_MIN ("svd-cutoff", :257-263).
 * n_registers > 1 builds that many independent n_qubits-wide registers in one handle (qubit index
 * = register*n_qubits + q); independent circuits then share batched kernel launches (config 4). */
int mps_create(int n_qubits, int n_registers, int max_bond, double svd_cutoff, int gauge, int device,
               uint64_t seed, mps_handle_t* out);
/* The MPI site-block scheme of the reference (process groups and block ownership ExaTnMpsVisitor.cpp:347-531, boundary 2q-gate
 * dispatch :2059-2170, finalize gather :685-696) as ONE process driving several GPUs of a box: the sites are split into
 * contiguous blocks, one per device (partition_by_cost != 0: blocks of equal estimated SVD cost on the saturated bond profile
 * of max_bond; otherwise equal counts -- one formula, the reference's two disagree when n % P != 0).  Each device runs its own
 * engine on its own stream and host thread; the gates of a dependency layer execute concurrently on all devices.  A gate on a
 * block boundary is executed by the owner of its LEFT site: the right owner's boundary tensor travels there and back as one
 * peer copy over NVLink each way (32 chi^2 bytes, ordered by CUDA events; no host staging, no collective).  Every other entry
 * point of this header works on such a handle unchanged (qubit indices are global).  n_devices == 1 is mps_create.  A device
 * may be listed more than once (its blocks then share that GPU: how the single-GPU tests exercise the exchange logic). */
int mps_create_sharded(int n_qubits, int max_bond, double svd_cutoff, int gauge, int n_devices, const int* devices,
                       int partition_by_cost, uint64_t seed, mps_handle_t* out);
/* the partition formula alone (no device needed): first_site[d] for d = 0..n_devices, first_site[n_devices] = n_qubits */
int mps_shard_partition(int n_qubits, int n_devices, int max_bond, int partition_by_cost, int* first_site);
/* the schedule of a flush on a sharded handle for a gate list (q1 = -1: single-qubit gate), as data, for tests of the host logic
 * (no device needed).  out receives records (device, kind: 0 LAYER / 1 SEND / 2 RECV, site, slot, n_gates, gate indices...),
 * devices in order, the ops of a device in execution order; *used = ints needed (the call fails when cap is smaller). */
int mps_shard_plan_debug(int n_qubits, int n_devices, const int* first_site, int count, const int* q0, const int* q1, int* out, int cap, int* used);
/* device blocks of a handle: *n_devices, and first_site[d] for d = 0..n_devices (first_site[n_devices] = n_qubits; may be NULL) */
int mps_shard_layout(mps_handle_t h, int* n_devices, int* first_site);
int mps_destroy(mps_handle_t h);
const char* mps_last_error(mps_handle_t h); /* h may be NULL: last mps_create error */
int mps_reset(mps_handle_t h);              /* back to |0...0>, ExaTnMpsVisitor.cpp:273-326 */
/* VQE mode (TNQVM.cpp:52-92, TNQVMVisitor::supportVqeMode): the ansatz is applied once, then every observable term runs its
 * change-of-basis gates on that state.  mps_snapshot remembers the current state of all registers on the device (sites, bond
 * spectra, discarded weight); mps_restore goes back to it.  The measure list is not part of the state (mps_clear_measure). */
int mps_snapshot(mps_handle_t h);
int mps_restore(mps_handle_t h);

/* keys: "max_bond", "svd_cutoff", "gauge" (as in mps_create), "cutoff_on_sqrt" (computePartialNormsSync ambiguity,
 * SURVEY 8c), "fuse_1q", "fuse_2q" (default 0: merge consecutive 2q gates on one site pair into one 4x4 -- fewer SVDs, but with
 * truncation active no longer truncation-for-truncation identical to the reference), "renormalize", "jacobi_tol", "null_tol", "jacobi_max_sweeps" (1..1000; matrices still rotating after that many sweeps are counted in mps_stats[11]), "profile",
 * "layer_batch" (0 = execute gate by gate), "norm_guard" (post-SVD check of ExaTnMpsVisitor.cpp:1632-1661: 2 = the reference's
 * Release-build behaviour (default): report once on stderr, count in mps_stats[12]; 1 = its Debug-build behaviour: the call fails;
 * 0 = off), "null_tol" (components with sigma <= null_tol * ||theta||_F are treated as the exact zeros they
 * stand for and dropped; <= 0 (default): 1e-13, independent of the matrix size and of what is batched together.  A documented deviation: LAPACK inside ExaTN keeps them as noise).  Options that change how queued gates execute flush the queue first.
 * Engine variants kept for A/B measurements, all parity-tested (tests/test_gpu_parity.py): "qr_prereduce" (default 1),
 * "jacobi_ctas_per_sm" (0 = by load, 1..4), "jacobi_wide_tasks" (1: eight warps per pair task when at most two tasks per SM
 * are resident), "jacobi_chunk_mb" (Jacobi work matrices of a layer run in chunks of at most this many MiB so that a chunk
 * stays L2-resident over its sweeps), "l2_persist" (persisting-L2 access window over the running chunk). */
int mps_set_option(mps_handle_t h, const char* key, double value);

/* applyGate 1q branch, ExaTnMpsVisitor.cpp:1185-1292.  m = row-major 2x2 complex. */
int mps_apply_1q(mps_handle_t h, int q, const double m[8]);
/* applyTwoQubitGate, ExaTnMpsVisitor.cpp:1387-1731 + truncateSvdTensors :2366-2536. m = row-major 4x4. */
int mps_apply_2q(mps_handle_t h, int q0, int q1, const double m[32]);
/* count independent (or not: dependencies are resolved) 2q gates in one call */
int mps_apply_layer(mps_handle_t h, int count, const int* q0, const int* q1, const double* mats);
/* a whole instruction list in one call (what TNQVM::execute's InstructionIterator loop delivers, TNQVM.cpp:126-134):
 * gate i acts on q0[i] (and q1[i], or -1 for a single-qubit gate); mats holds 32 doubles per gate, a row-major 4x4 complex
 * matrix or a row-major 2x2 in the first 8 */
int mps_apply_gates(mps_handle_t h, int count, const int* q0, const int* q1, const double* mats);
/* gates are queued and executed in dependency layers; flush forces execution, sync also waits */
int mps_flush(mps_handle_t h);
int mps_sync(mps_handle_t h);

/* "norm" extra-info, ExaTnMpsVisitor.cpp:604-612 (<psi|psi>, by transfer-matrix sweep) */
int mps_norm(mps_handle_t h, int reg, double* out);
/* "exp-val-z", ExaTnMpsVisitor.cpp:616-644: <psi| prod Z_q |psi>, NOT divided by the norm */
int mps_expval_z(mps_handle_t h, int reg, int nq, const int* qubits, double* out);
/* <Z_k> for every qubit of a register (left/right environment sweeps; ITensorMPSVisitor.cpp:173-250) */
int mps_expval_z_all(mps_handle_t h, int reg, double* out_n);
/* <Z_i Z_j> for a list of pairs */
int mps_expval_zz_pairs(mps_handle_t h, int reg, int npairs, const int* qi, const int* qj, double* out);
/* computeWaveFuncSlice, ExaTnMpsVisitor.cpp:2588-2675: bits[k] in {0,1} fixed or -1 = open leg.
 * out receives 2^(#open) complex amplitudes (open qubits ordered by index, lowest = fastest). */
int mps_amplitude(mps_handle_t h, int reg, const int8_t* bits, double* out, size_t* len);
/* full state vector, evaluateSync(ket) at ExaTnMpsVisitor.cpp:591-597 (n_qubits <= 30) */
int mps_statevector(mps_handle_t h, int reg, double* out);

/* visit(Measure), ExaTnMpsVisitor.cpp:991-994: records the qubit; character i of a sample string
 * belongs to the i-th recorded qubit */
int mps_measure(mps_handle_t h, int q);
int mps_clear_measure(mps_handle_t h);
int mps_seed(mps_handle_t h, uint64_t seed); /* {"seed", int}, TNQVM.hpp:114-117 */
int mps_n_measured(mps_handle_t h, int* out); /* qubits recorded by mps_measure since the last reset / clear (repeats count) */
/* finalize() sampling: n < 20 -> GenerateSamples on the state vector (GateMatrixAlgebra.hpp:125-156),
 * n >= 20 -> per-shot sequential RDM sampling (getMeasureSample, ExaTnMpsVisitor.cpp:2211-2364).
 * out: shots * n_measured chars (no terminators), out_cap = its capacity in chars (the call fails when it is smaller than
 * shots * n_measured); n_out = strings actually produced. */
int mps_sample(mps_handle_t h, int reg, int shots, char* out, size_t out_cap, int* n_out);

int mps_bond_dims(mps_handle_t h, int* out);                                           /* n_total-1 entries */
int mps_singular_values(mps_handle_t h, int bond, double* out, int cap, int* count);   /* of the last SVD on that bond */
int mps_discarded_weight(mps_handle_t h, double* out); /* sum over truncations of w = discarded/total weight */
/* fidelity estimate of a truncated run, prod over truncations of (1 - w) (BASELINE config 5; the reference keeps no such record,
 * ExaTnMpsVisitor.cpp:2366-2536) */
int mps_fidelity_estimate(mps_handle_t h, double* out);
int mps_get_site(mps_handle_t h, int k, double* out, int shape[3]);                    /* out may be NULL (shape only) */
int mps_set_site(mps_handle_t h, int k, const double* in, int dl, int dr);
/* the same for a list of sites with ONE wait at the end (upload / download of a host-resident state: 2 calls instead of 2 n);
 * out[i] must hold the 2 * dl * dr complex numbers of site k[i] (shapes from mps_get_site with out = NULL, or mps_bond_dims) */
int mps_set_sites(mps_handle_t h, int count, const int* k, const double* const* in, const int* dl, const int* dr);
int mps_get_sites(mps_handle_t h, int count, const int* k, double* const* out);

/* site-sharded multi-GPU support (replaces replicateTensorSync at ExaTnMpsVisitor.cpp:2088-2158):
 * device pointer of a site tensor for NCCL send/recv, and adoption of a received tensor */
int mps_site_device_ptr(mps_handle_t h, int k, void** dptr, int shape[3]);
int mps_resize_site(mps_handle_t h, int k, int dl, int dr, void** dptr);

/* counters: [0] 2q gates executed, [1] 1q kernel gates, [2] layers, [3] jacobi sweeps, [4] kernel launches,
 * [5] ms merge GEMM, [6] ms SVD (QR pre-reduction + Jacobi), [7] ms truncate+write-back, [8] ms of [6] spent in the QR
 * pre-reduction (5..8 only with option "profile"), [9] real flops issued on the DMMA pipe by the Jacobi pair tasks (process-wide),
 * [10] 2q gates merged into their predecessor on the same site pair (option "fuse_2q"), [11] SVDs that had not converged
 * after "jacobi_max_sweeps" sweeps, [12] norm-guard violations, [13] boundary exchanges and [14] bytes moved between devices
 * (site-sharded handles).  On a sharded handle the counters are sums over the devices. */
int mps_stats(mps_handle_t h, double* out, int cap);
/* the CUDA stream (cudaStream_t) all work of this handle is issued on: callers time with events recorded on it
 * and order their own transfers (NCCL send/recv of boundary sites) against it */
int mps_get_stream(mps_handle_t h, void** stream);

#ifdef __cplusplus
}
#endif
#endif
