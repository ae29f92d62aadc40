/* gpuamr_b200.h — C ABI of the B200-native gpu-amr hot path.
 *
 * Everything the host side (C++ templates in include/ndtree, include/solver; the ctypes
 * harness in gpu-amr_b200/binding.py) needs from the GPU goes through these entry points:
 * plain pointers, sizes and POD structs, no C++ or torch types.  Each group cites the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns amrb_status (0 = AMRB_OK); amrb_last_error() gives the text of
 *     the last failure on the calling thread.  The reference throws std::runtime_error across
 *     this boundary (src/cuda/halo_exchange.cu:16-26) or silently drops launch errors
 *     (src/cuda/fvm_time_step.cu:142,238,244,261,282); the inline C++ shim re-throws.
 *   - all launches are asynchronous on the pool's stream unless stated otherwise.
 *   - patches are padded row-major tensors, last layout dim fastest, every dim padded by
 *     2*halo (containers/static_layout.hpp:28-37, ndtree/patch_layout.hpp:14-116); field f of
 *     patch p lives at field_base[f] + p*flat_size (ndtree.hpp:166-174).
 *   - directions: d = 2*layout_dim + (positive ? 1 : 0)  (ndtree/neighbor.hpp:103-255).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GPUAMR_B200_H
#define GPUAMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int amrb_status;
enum {
    AMRB_OK            = 0,
    AMRB_ERR_ARGUMENT  = 1,
    AMRB_ERR_CUDA      = 2,
    AMRB_ERR_CAPACITY  = 3,
    AMRB_ERR_UNSUPPORTED = 4, /* shape/equation combination not instantiated */
    AMRB_ERR_STATE     = 5
};

enum { AMRB_EQ_ADVECTION = 0, AMRB_EQ_EULER = 1 };             /* solver/{Advection,Euler}Physics.hpp */
enum { AMRB_REL_NONE = 0, AMRB_REL_SAME = 1, AMRB_REL_FINER = 2, AMRB_REL_COARSER = 3 };
                                                                /* cuda/halo_exchange.hpp:11-17 */
enum { AMRB_STABLE = 0, AMRB_REFINE = 1, AMRB_COARSEN = 2 };   /* ndtree.hpp:265-270 refine_status_t */
/* device storage of a field-patch:
 *   PADDED   the reference's device layout (ndtree.hpp:341-362): prod(size + 2*halo) doubles, ghosts stored;
 *   INTERIOR prod(size) doubles, no ghosts anywhere in HBM: the fused step gathers its ghost layer from the
 *            neighbor interiors anyway, so storing ghosts only costs traffic and memory (3D 8^3 patches:
 *            6 400 + 6 400 B moved per 4 096 algorithmic; 1.07e9 cells: 168 GB instead of 86 GB).  Rank 3,
 *            halo 1, sizes 8 and 16.  The host exchange format (amrb_pool_upload / download, VTK output)
 *            stays the padded patch; the padded image is built in a staging buffer on demand. */
enum { AMRB_STORAGE_PADDED = 0, AMRB_STORAGE_INTERIOR = 1 };

const char* amrb_last_error(void);
const char* amrb_version(void);
/* number of visible CUDA devices (0 on a CPU-only box); never fails */
int amrb_device_count(void);

/* ------------------------------------------------------------------------------------------
 * 1. memory & sync — replaces amr::cuda::device_malloc … async_copy_fence_wait
 *    (include/cuda/device_buffer.hpp:9-48, src/cuda/device_buffer.cu:31-239)
 * ---------------------------------------------------------------------------------------- */
amrb_status amrb_device_malloc(void** out, size_t bytes);
amrb_status amrb_device_free(void* ptr);
amrb_status amrb_host_pinned_malloc(void** out, size_t bytes);
amrb_status amrb_host_pinned_free(void* ptr);
amrb_status amrb_copy_host_to_device(void* dst, const void* src, size_t bytes);
amrb_status amrb_copy_host_to_device_async(void* dst, const void* src, size_t bytes, void* stream);
amrb_status amrb_copy_device_to_host(void* dst, const void* src, size_t bytes);
amrb_status amrb_copy_device_to_host_async(void* dst, const void* src, size_t bytes, void* stream);
amrb_status amrb_copy_device_to_device(void* dst, const void* src, size_t bytes);
amrb_status amrb_stream_create(void** out);          /* async_copy_stream_create  */
amrb_status amrb_stream_destroy(void* stream);
amrb_status amrb_stream_synchronize(void* stream);
amrb_status amrb_stream_wait_fence(void* stream, void* fence);
amrb_status amrb_fence_create(void** out);           /* async_copy_fence_create   */
amrb_status amrb_fence_destroy(void* fence);
amrb_status amrb_fence_record(void* fence, void* stream);
amrb_status amrb_fence_wait(void* fence);
amrb_status amrb_device_synchronize(void);

/* ------------------------------------------------------------------------------------------
 * 2. patch layout (ndtree/patch_layout.hpp:14-116, containers/container_utils.hpp:56-64)
 * ---------------------------------------------------------------------------------------- */
typedef struct amrb_layout
{
    int32_t rank;      /* 2 | 3 */
    int32_t size[3];   /* interior cells per layout dim, dim 0 slowest; unused = 1 */
    int32_t halo;      /* ghost width per side (1 | 2) */
    int32_t nvar;      /* fields per cell: 1 (advection) or rank+2 (Euler) */
    int32_t equation;  /* AMRB_EQ_* */
    int32_t depth;     /* morton_id<Depth,Rank>: finest level */
    int32_t storage;   /* AMRB_STORAGE_* */
} amrb_layout;

/* 1 if fused kernels for this shape were compiled into the library */
int    amrb_layout_supported(const amrb_layout* layout);
size_t amrb_layout_flat_size(const amrb_layout* layout); /* prod(size+2*halo): the padded (host) patch */
size_t amrb_layout_data_size(const amrb_layout* layout); /* prod(size)        */
size_t amrb_layout_storage_size(const amrb_layout* layout); /* doubles per field-patch in the pool */

/* ------------------------------------------------------------------------------------------
 * 3. device patch pool — replaces ndtree's m_data_buffers / m_next_buffers device mirrors,
 *    halo metadata and level buffers (ndtree.hpp:341-409, 561-679, 1606-1700, 2004-2060)
 * ---------------------------------------------------------------------------------------- */
typedef struct amrb_pool amrb_pool;

/* device-resident SoA pool of `capacity` patches, current + next buffer per field,
 * zero-initialised (corner ghosts are observable by max criteria). */
amrb_status amrb_pool_create(const amrb_layout* layout, size_t capacity, int device,
                             amrb_pool** out);
/* same, over caller-owned device memory (e.g. torch tensors): cur[f], nxt[f] hold
 * capacity*flat_size doubles each; stream may be NULL (legacy default stream). */
amrb_status amrb_pool_create_external(const amrb_layout* layout, size_t capacity, int device,
                                      double* const* cur, double* const* nxt, void* stream,
                                      amrb_pool** out);
amrb_status amrb_pool_destroy(amrb_pool* pool);
size_t      amrb_pool_capacity(const amrb_pool* pool);
size_t      amrb_pool_size(const amrb_pool* pool);
void*       amrb_pool_stream(const amrb_pool* pool);
/* current-buffer device pointer of field f (ndtree::get_device_buffer, ndtree.hpp:561-581); patch p
 * starts at p * amrb_layout_storage_size(layout) doubles */
double*     amrb_pool_field(const amrb_pool* pool, int field);
double*     amrb_pool_next_field(const amrb_pool* pool, int field);
/* current <-> next (ndtree::swap_buffers, ndtree.hpp:1558-1579) for callers that filled the next buffer
 * themselves, e.g. patch migration between the Morton ranges of a sharded mesh */
amrb_status amrb_pool_swap_buffers(amrb_pool* pool);

/* neighbor / halo index tables in the reference's host form, one row per (patch, direction):
 *   levels[P]; rel[P][2R]; nbr[P][2R][2^(R-1)] (linear patch index, -1 padded);
 *   quad[P][2R][R] contact quadrant of a coarser neighbor (ndtree/neighbor.hpp:22-100,
 *   cuda/halo_exchange.hpp:19-26, ndtree.hpp:1606-1660 rebuild_halo_exchange_metadata).
 * n_total >= n_owned: slots [n_owned, n_total) are ghost slots (copies of patches owned by
 * another GPU) that are only ever read as halo sources.  Converts to the compact device
 * format and uploads; call once after every refine/coarsen. */
amrb_status amrb_pool_set_topology(amrb_pool* pool, size_t n_owned, size_t n_total,
                                   const int32_t* levels, const int8_t* rel, const int32_t* nbr,
                                   const int8_t* quad);
/* the same tables built ON THE DEVICE from the ascending leaf ids (id = morton << 6 | level) of the whole
 * mesh — replaces the host loop of ndtree::rebuild_halo_exchange_metadata (ndtree.hpp:1606-1660) and its
 * H2D copy of the metadata by one copy of 8 bytes per leaf and one kernel.  Single-GPU
 * form (no ghost slots).  Bit-identical to amrb_tree_tables + amrb_pool_set_topology. */
amrb_status amrb_pool_set_topology_from_ids(amrb_pool* pool, const uint64_t* ids, size_t n);
/* device tables read back in the compact device form: levels[n], meta[n][2R] = relation | contact-quadrant
 * bits << 2, nbr[n][2R][2^(R-1)] (tests, debugging) */
amrb_status amrb_pool_get_tables(amrb_pool* pool, int32_t* levels, uint8_t* meta, int32_t* nbr);
/* physical domain lengths per *physical* axis (x,y,z) and solver constants
 * (solver/physics_system.hpp:58-85, amr_solver.hpp:62-69) */
amrb_status amrb_pool_set_physics(amrb_pool* pool, const double* lengths, double gamma,
                                  double cfl);

/* bulk host<->device transfer of whole padded patches of the CURRENT buffer
 * (ndtree::sync_current_{to,from}_device, ndtree.hpp:583-640); blocking */
amrb_status amrb_pool_upload(amrb_pool* pool, int field, size_t first_patch, size_t n_patches,
                             const double* host);
amrb_status amrb_pool_download(amrb_pool* pool, int field, size_t first_patch, size_t n_patches,
                               double* host);
/* the same into the NEXT buffer, for host-side steppers that fill ndtree::get_out_patch<Map, next_buffer>
 * and then call swap_buffers (ndtree.hpp:516-559, 1558-1579); blocking */
amrb_status amrb_pool_upload_next(amrb_pool* pool, int field, size_t first_patch, size_t n_patches,
                                  const double* host);
/* interior-only variants (host array is [n_patches][prod(size)]) */
amrb_status amrb_pool_upload_interior(amrb_pool* pool, int field, size_t first_patch,
                                      size_t n_patches, const double* host);
amrb_status amrb_pool_download_interior(amrb_pool* pool, int field, size_t first_patch,
                                        size_t n_patches, double* host);

/* ------------------------------------------------------------------------------------------
 * 4. halo fill — replaces halo_exchange_scalar_patches_inplace, one launch per field
 *    (include/cuda/halo_exchange.hpp:42-47, src/cuda/halo_exchange.cu:291-370) and the CPU
 *    operators same_t / finer_t / coarser_t (ndtree/patch_utils.hpp:303-441).
 *    One launch materialises the face halos of ALL fields of the current buffer, in place.
 * ---------------------------------------------------------------------------------------- */
amrb_status amrb_pool_halo_exchange(amrb_pool* pool);
/* lazy materialisation of the face halos: with `on`, amrb_pool_advance_batch_async no longer ends with a
 * halo fill (post-condition of amr_solver::time_step, amr_solver.hpp:351-352); the fill happens when the
 * padded patches are observed (amrb_pool_download, amrb_pool_patch_max_flags, amrb_pool_apply_plan) or
 * on request (amrb_pool_ensure_halos — call it before handing out amrb_pool_field pointers). */
amrb_status amrb_pool_set_lazy_halos(amrb_pool* pool, int on);
amrb_status amrb_pool_ensure_halos(amrb_pool* pool);

/* ------------------------------------------------------------------------------------------
 * 5. time stepping — replaces launch_compute_dt_kernel_device, launch_finalize_step_dt,
 *    launch_time_step_kernel_with_device_dt, launch_set_{double,uint32}_buffer
 *    (include/cuda/fvm_time_step.hpp:10-58) and the loop in
 *    amr_solver::advance_batch_async / finish_advance_batch (solver/amr_solver.hpp:155-262).
 * ---------------------------------------------------------------------------------------- */
/* cfl * min dx/speed over the current buffer (amr_solver.hpp:355-413); blocking */
amrb_status amrb_pool_compute_dt(amrb_pool* pool, double* dt_out);
/* one explicit step with a host-provided dt: fused halo gather + flux + update into the
 * next buffer, then swap (amr_solver::time_step_cpu, amr_solver.hpp:265-353).  Global face
 * halos of the new current buffer are NOT materialised (call amrb_pool_halo_exchange). */
amrb_status amrb_pool_step(amrb_pool* pool, double dt);
/* up to `steps` steps with device-resident dt / remaining-time / step-count scalars, each
 * step one fused launch (halo gather + flux + update + next-dt reduction).  After the last
 * step face halos are materialised, so the post-condition of the reference holds:
 * current buffers hold the new state with face halos filled.  Asynchronous. */
amrb_status amrb_pool_advance_batch_async(amrb_pool* pool, size_t steps, double remaining_time);
/* waits for the batch; returns sum of step dts, number of executed (dt > 0) steps and,
 * if dts != NULL, the first min(executed, dts_capacity) individual step sizes */
amrb_status amrb_pool_finish_advance_batch(amrb_pool* pool, double* dt_sum, size_t* executed,
                                           double* dts, size_t dts_capacity);
/* statistics for bench accounting: kernels launched by this pool since creation */
uint64_t amrb_pool_launch_count(const amrb_pool* pool);
/* select the step implementation: 0 = fused lazy-halo pipeline (default),
 * 1 = unfused (materialise halos every step, then the same stencil kernel without gather),
 * 2 = first-generation fused kernel (thread per cell) — 1 and 2 are kept for A/B measurement */
amrb_status amrb_pool_set_mode(amrb_pool* pool, int mode);
/* kernel variant inside mode 0, for A/B measurement and variant-agreement tests (0 = default; the values
 * are listed next to the dispatch in csrc/amrb_api.cu; initial value = environment AMRB_VARIANT) */
amrb_status amrb_pool_set_variant(amrb_pool* pool, int variant);
/* The carried CFL minimum: a batch ends with the dt-min of its final state, and the next batch starts
 * from it without a compute_dt pass.  Every entry point of this library that changes the state
 * (upload, apply_plan, set_topology, swap_buffers, set_physics) drops it.  A caller that writes the
 * current buffer through amrb_pool_field / amrb_pool_next_field pointers (get_device_buffer) MUST call
 * amrb_pool_mark_dirty afterwards, otherwise the next batch takes its first step with the dt of the
 * overwritten state. */
amrb_status amrb_pool_mark_dirty(amrb_pool* pool);

/* the same batch, decomposed so that a multi-GPU driver can interleave ghost-face traffic:
 *   batch_begin(max_steps, remaining)
 *   per step: step_partial(list A) ... step_partial(list B); [all-reduce(min) of dtmin_slot(k+1)];
 *             step_commit()                      (swap buffers, k -> k+1)
 *   batch_end()                                  (materialise halos, start the scalar read-back)
 * dev_list = device array of owned patch indices (NULL = all owned patches).  Every partial
 * launch of step k reads the same dt scalars and min-reduces into slot k+1. */
amrb_status amrb_pool_batch_begin(amrb_pool* pool, size_t max_steps, double remaining_time);
amrb_status amrb_pool_step_partial(amrb_pool* pool, const int32_t* dev_list, size_t count);
amrb_status amrb_pool_step_commit(amrb_pool* pool);
amrb_status amrb_pool_batch_end(amrb_pool* pool, int materialise_halos);
/* device address of the dt-min slot entering step k of the current batch (a positive double;
 * min across GPUs = all-reduce(min) on it as float64) */
double*     amrb_pool_dtmin_slot(amrb_pool* pool, size_t k);

/* ------------------------------------------------------------------------------------------
 * 6. inter-GPU ghost faces (no reference counterpart: the reference is single-GPU).
 *    pack: gathers, for each (patch, direction) entry, the min(2h, S)-thick interior slab next to
 *    that face of every field (covers the same / coarser / finer halo operators) into a contiguous
 *    send buffer; unpack: scatters a received buffer into the same slab of a ghost slot.
 *    Entry = {int32 patch, int32 direction | layers << 4}; buffer = [entry][field][layer][face cell].
 *    layers = how many of the slab's layers are moved (0 = all min(2h, S)); h are enough for a face no
 *    COARSER patch reads (same-level copy and injection into a finer patch use the first h layers only).
 *    The slab keeps its fixed size and layout either way.
 * ---------------------------------------------------------------------------------------- */
size_t      amrb_pool_face_slab_doubles(const amrb_pool* pool, int direction); /* per field */
amrb_status amrb_pool_pack_faces(amrb_pool* pool, const int32_t* dev_entries, size_t count,
                                 double* dev_buffer);
amrb_status amrb_pool_unpack_faces(amrb_pool* pool, const int32_t* dev_entries, size_t count,
                                   const double* dev_buffer);

/* ------------------------------------------------------------------------------------------
 * 6b. the same exchange over peer-mapped memory (NVLink / NVSwitch), no collective library on the data
 *    path: per step ONE kernel gathers the slabs, stores them straight into the peers' receive buffers,
 *    drops this rank's CFL minimum into the peers' mailboxes and raises its arrival flag there; the
 *    receiver acquire-spins on the flags (one warp, bounded by a timeout), folds the minima (= the
 *    all-reduce(min) of dt) and unpacks locally.  The K-step loop runs inside the library.
 *    Setup: every rank creates its exchange from its ShardPlan lists, exports its mailbox and its two
 *    receive buffers (amrb_exchange_buffer + amrb_ipc_export), opens the peers' handles (amrb_ipc_open; in
 *    one process: the peers' pointers as they are) and connects them.  world <= 8.
 *      send_entries [n_send][2] = {owned patch, face}, sorted by destination rank (send_counts[world]);
 *      send_offsets[r] = entries ranks below this one send to r (where this rank's segment starts there);
 *      recv_entries [n_recv][2] = {ghost slot, face} in arrival order (source rank, then the sender's order).
 * ---------------------------------------------------------------------------------------- */
typedef struct amrb_exchange amrb_exchange;
amrb_status amrb_exchange_create(amrb_pool* pool, int rank, int world, const int32_t* send_entries,
                                 const int64_t* send_counts, const int64_t* send_offsets,
                                 const int32_t* recv_entries, size_t n_recv, amrb_exchange** out);
amrb_status amrb_exchange_destroy(amrb_exchange* ex);
/* which: 0 = mailbox, 1 / 2 = receive buffer of generation parity 0 / 1 (device pointers of THIS rank) */
void*       amrb_exchange_buffer(amrb_exchange* ex, int which);
amrb_status amrb_ipc_export(void* dev_ptr, void* handle64);        /* cudaIpcGetMemHandle: 64 bytes */
amrb_status amrb_ipc_open(const void* handle64, void** dev_ptr);   /* maps a peer process's buffer  */
amrb_status amrb_ipc_close(void* dev_ptr);
amrb_status amrb_exchange_connect(amrb_exchange* ex, int peer, void* mailbox, void* recv0, void* recv1);
/* one exchange of the current buffers' slabs into the peers' ghost slots; asynchronous */
amrb_status amrb_exchange_halo(amrb_exchange* ex);
/* amrb_pool_advance_batch_async over the sharded mesh: per step push + wait/fold + unpack + fused step;
 * ends with an exchange and the halo materialisation.  Read back with amrb_pool_finish_advance_batch. */
amrb_status amrb_exchange_advance_batch_async(amrb_exchange* ex, size_t steps, double remaining_time,
                                              int overlap);
/* owned patches with / without a remote neighbor: with them, overlap = 1 above runs the boundary patches first
 * and pushes their slabs from the next buffer on a side stream while the interior patches are advanced */
amrb_status amrb_exchange_set_lists(amrb_exchange* ex, const int32_t* boundary, size_t n_boundary,
                                    const int32_t* interior, size_t n_interior);
/* the two halves of one exchange (drivers that interleave several ranks in one process); with_dt folds the
 * CFL minimum of slot k of the open batch */
amrb_status amrb_exchange_push(amrb_exchange* ex, int with_dt, size_t k);
amrb_status amrb_exchange_wait(amrb_exchange* ex, int with_dt, size_t k);
/* per-phase device times [ms] of the last batch (plain schedule): out4 = push, wait, unpack, fused step */
amrb_status amrb_exchange_set_timing(amrb_exchange* ex, int on);
amrb_status amrb_exchange_get_timing(amrb_exchange* ex, double* out4);
/* 1 when a receiver gave up waiting for a peer since the last call (state undefined afterwards) */
int         amrb_exchange_timed_out(amrb_exchange* ex);
uint64_t    amrb_exchange_launch_count(const amrb_exchange* ex);

/* ------------------------------------------------------------------------------------------
 * 7. host topology — Morton-ordered leaf set, refine/coarsen with 2:1 balancing and the
 *    neighbor tables derived from it.  Replaces the host side of ndtree::reconstruct_tree
 *    (ndtree.hpp:886-940, 1127-1271) and neighbor maintenance (ndtree/neighbor.hpp:291-572)
 *    with a set-based formulation: tables depend only on the set of leaves.
 * ---------------------------------------------------------------------------------------- */
typedef struct amrb_tree amrb_tree;

amrb_status amrb_tree_create(int rank, int depth, amrb_tree** out); /* single periodic root */
amrb_status amrb_tree_destroy(amrb_tree* tree);
size_t      amrb_tree_size(const amrb_tree* tree);
const uint64_t* amrb_tree_ids(const amrb_tree* tree); /* ascending morton ids (<<6 | level) */
/* one reconstruct pass. flags[i] in AMRB_{STABLE,REFINE,COARSEN} for leaf i.  On return
 * `plan` (if not NULL, capacity >= new size) holds for each new leaf the transfer source:
 * plan_kind[i] 0 = copy of old leaf plan_src[i]; 1 = prolongation from old leaf plan_src[i]
 * (child number in plan_child[i]); 2 = restriction of old leaves plan_src[i] .. +2^rank-1.
 * returns changed = 1 when the leaf set changed. */
amrb_status amrb_tree_reconstruct(amrb_tree* tree, const int8_t* flags, size_t capacity,
                                  int* changed);
size_t      amrb_tree_plan_size(const amrb_tree* tree);
amrb_status amrb_tree_plan(const amrb_tree* tree, int8_t* kind, int32_t* src, int8_t* child);
/* neighbor tables of the current leaf set (same format as amrb_pool_set_topology) */
amrb_status amrb_tree_tables(const amrb_tree* tree, int32_t* levels, int8_t* rel, int32_t* nbr,
                             int8_t* quad);
uint64_t amrb_morton_encode(int rank, const uint32_t* coords, int level);
void     amrb_morton_decode(int rank, uint64_t id, uint32_t* coords, int* level);

/* apply a reconstruct plan on the device: gathers the new current buffer from the old one
 * (copy / prolongation / restriction fused with the Morton re-sort), replacing
 * interpolate/restrict_scalar_patches_inplace + permute_patches_inplace_batch
 * (include/cuda/intergrid_transfer.hpp:28-42, include/cuda/permutation.hpp:10-17). */
amrb_status amrb_pool_apply_plan(amrb_pool* pool, size_t new_size, const int8_t* kind,
                                 const int32_t* src, const int8_t* child);

/* per-patch refinement decisions on the device: max over ALL flat cells of one field
 * (benchmark criterion, src/cuda/fvm_refinement_criterion.cu:27-67):
 * Refine if max > refine_threshold && level < max_level; Coarsen if max < coarsen_threshold
 * && level > min_level; else Stable.  Blocking; flags is a host array [size]. */
amrb_status amrb_pool_patch_max_flags(amrb_pool* pool, int field, double refine_threshold,
                                      double coarsen_threshold, int min_level, int max_level,
                                      int8_t* flags);

/* the same criterion with the flags LEFT ON THE DEVICE (no read-back; input of amrb_pool_reconstruct_device) */
amrb_status amrb_pool_flag_patches(amrb_pool* pool, int field, double refine_threshold,
                                   double coarsen_threshold, int min_level, int max_level);
/* reconstruct_tree on the device (replaces the host side of ndtree.hpp:886-940, 1127-1271):
 * selection (eligibility, 2:1 ripple, coarsening veto) over the device halo tables, new leaf ids in Morton
 * order, transfer plan, data motion and new halo tables -- the leaf set and the flags never leave the GPU; the
 * host reads back two integers.  dev_flags: device array [size] of AMRB_{STABLE,REFINE,COARSEN} (NULL = the
 * flags of the last amrb_pool_flag_patches).  Needs the topology built by amrb_pool_set_topology_from_ids
 * (single GPU, no ghost slots).  Bit-identical to amrb_tree_reconstruct + amrb_pool_apply_plan +
 * amrb_pool_set_topology_from_ids.  Blocking (one 8-byte read-back). */
amrb_status amrb_pool_reconstruct_device(amrb_pool* pool, const int8_t* dev_flags, int* changed,
                                         size_t* new_size);
/* leaf ids of the device topology (ascending), to mirror a device reconstruct into an amrb_tree */
amrb_status amrb_pool_get_ids(amrb_pool* pool, uint64_t* host_ids, size_t capacity);
amrb_status amrb_tree_assign(amrb_tree* tree, const uint64_t* ids, size_t n);
/* transfer plan of the last device reconstruct: kind / src / child per new leaf (tests) */
amrb_status amrb_pool_get_plan(amrb_pool* pool, int8_t* kind, int32_t* src, int8_t* child);

/* raw-pointer form of the same criterion — replaces
 * amr::cuda::compute_scalar_patch_amr_decisions_from_device
 * (include/cuda/fvm_refinement_criterion.hpp:20-27): all pointers are device pointers,
 * decisions[i] in AMRB_{STABLE,REFINE,COARSEN}; asynchronous on `stream` (NULL = default). */
amrb_status amrb_patch_max_flags_device(const double* dev_field, const int32_t* dev_levels,
                                        size_t num_patches, size_t cells_per_patch,
                                        double refine_threshold, double coarsen_threshold,
                                        int min_level, int max_level, int8_t* dev_decisions,
                                        void* stream);
/* device pointer of the per-patch level array uploaded by amrb_pool_set_topology
 * (ndtree::get_device_patch_level_buffer, ndtree.hpp:659-678) */
const int32_t* amrb_pool_levels(const amrb_pool* pool);

/* ------------------------------------------------------------------------------------------
 * 9. kernel-level entry points on CALLER-OWNED padded device arrays — the reference's own launch protocol
 *    (one call per kernel of src/cuda/halo_exchange.cu and src/cuda/fvm_time_step.cu), for a build that
 *    keeps the reference's ndtree.hpp / amr_solver.hpp and replaces only those two translation units
 *    (integration/amrb_shim.cpp).  Arrays use the reference's device layout
 *    (field f of patch p at base[f] + p * flat_size); all launches are asynchronous on `stream`
 *    (NULL = the legacy default stream the reference uses).
 *      amrb_raw_halo_exchange   halo_exchange_scalar_patches_inplace (include/cuda/halo_exchange.hpp:42-47):
 *                               ONE field in place; ref_metadata = device array of the reference's 36-byte
 *                               halo_direction_metadata records, [num_patches * 2 * rank]
 *      amrb_raw_compute_dt      launch_compute_dt_kernel_device (fvm_time_step.hpp): *dev_dt = min dx / speed
 *      amrb_raw_finalize_dt     launch_finalize_step_dt: dt *= cfl, clamp to the remaining time, accumulate
 *      amrb_raw_time_step       launch_time_step_kernel_with_device_dt: in -> out with the step size at
 *                               *dev_dt, stored face ghosts of `in` trusted (the reference fills them first)
 *    root_cell_size[d] = cell size of a level-0 patch along solver direction d (x, y, z)
 *    (time_step_launch_config::root_c_size). */
amrb_status amrb_raw_halo_exchange(const amrb_layout* layout, double* field_base, const void* ref_metadata,
                                   size_t metadata_count, size_t num_patches, void* stream);
amrb_status amrb_raw_compute_dt(const amrb_layout* layout, const double* const* fields, const int32_t* levels,
                                size_t num_patches, const double* root_cell_size, double gamma,
                                double* dev_dt, void* stream);
amrb_status amrb_raw_finalize_dt(double* dev_dt, double* dev_accumulator, double* dev_remaining,
                                 uint32_t* dev_step_count, double cfl, void* stream);
amrb_status amrb_raw_time_step(const amrb_layout* layout, double* const* in, double* const* out,
                               const int32_t* levels, size_t num_patches, const double* root_cell_size,
                               double gamma, const double* dev_dt, void* stream);
amrb_status amrb_raw_set_double(double* dev, double value, void* stream);
amrb_status amrb_raw_set_uint32(uint32_t* dev, uint32_t value, void* stream);

/* ------------------------------------------------------------------------------------------
 * 8. profiling hooks — replaces amr::cuda::profile_capture_{start,stop}, profile_range_{push,pop}
 *    (include/cuda/profiler.hpp, src/cuda/device_buffer.cu:229-237)
 * ---------------------------------------------------------------------------------------- */
amrb_status amrb_profile_capture_start(void);
amrb_status amrb_profile_capture_stop(void);
amrb_status amrb_profile_range_push(const char* label);
amrb_status amrb_profile_range_pop(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUAMR_B200_H */
