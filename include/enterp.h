/*
 * enterp.h — C ABI of the B200-native batched curve-evaluation engine for enterpolation 0.3.0.
 *
 * The reference (/root/reference, 100 % Rust) has no FFI of its own; its "operator API" for the hot
 * path is the trait surface
 *     Signal<R>::eval / extract / sample   src/base/signal.rs:15-19, 48-58, 166-173
 *     Curve<R>::domain / take              src/base/signal.rs:194, 221-228
 * reached through the typestate builders
 *     Linear::builder()                    src/linear/mod.rs:105      (src/linear/builder.rs)
 *     Bezier::builder()                    src/bezier/mod.rs:207      (src/bezier/builder.rs)
 *     BSpline::builder()                   src/bspline/mod.rs:115     (src/bspline/builder.rs)
 *     Weighted<..> (NURBS)                 src/weights/weighted.rs:28-37
 * Every entry point below states which of those it stands in for. A Rust `-sys` crate would bind
 * exactly these symbols; in this repo the callers are the C++ fluent mirror
 * (include/enterpolation.hpp) and the Python ctypes mirror (enterpolation_b200/).
 *
 * Conventions
 *  - plain C, no torch / CUDA types in signatures: device pointers are `void*`, streams are
 *    `void*` (a cudaStream_t, NULL = legacy default stream);
 *  - the caller owns every host array (copied during create: "upload once") and every device
 *    buffer; the library owns only the opaque handle;
 *  - a handle is immutable after create; concurrent calls from several host threads are allowed
 *    (mirrors `eval(&self)` re-entrancy, src/base/space.rs:63-66);
 *  - outputs are dense AoS `[n][dim]` of the curve scalar (weighted curves emit the `dim`
 *    projected components, src/weights/homogeneous.rs:134-136);
 *  - where the reference would panic (NaN/inf/negative into `to_usize().unwrap()`,
 *    src/base/list.rs:456,486,539-540; a B-spline span index below the lower cut — descending or
 *    collapsed equidistant knots — at the element access of src/bspline/mod.rs:131) the device writes
 *    NaN into every component of that sample and bumps a per-curve counter readable with
 *    enterp_invalid_count();
 *  - no CPU fallback exists: every eval entry point returns ENTERP_NO_DEVICE / ENTERP_CUDA_ERROR
 *    when no B200 is usable.
 */
#ifndef ENTERP_H_
#define ENTERP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ENTERP_API
#else
#define ENTERP_API __attribute__((visibility("default")))
#endif

typedef struct enterp_curve enterp_curve; /* opaque; immutable after create; owns device replicas */

/* Error taxonomy = union of LinearError (src/linear/error.rs:13-20), BezierError
 * (src/bezier/error.rs:12-17) and BSplineError (src/bspline/error.rs:16-31), plus ABI-level codes. */
typedef enum enterp_status {
  ENTERP_OK = 0,
  ENTERP_TOO_FEW_ELEMENTS = 1,            /* builder.rs:107  */
  ENTERP_TOO_FEW_KNOTS = 2,               /* builder.rs:135  */
  ENTERP_KNOT_ELEMENT_INEQUALITY = 3,     /* linear/error.rs:56 */
  ENTERP_NOT_SORTED = 4,                  /* base/list.rs:325 */
  ENTERP_INCONGRUOUS_ELEMENTS_KNOTS = 5,  /* bspline/error.rs:132 */
  ENTERP_INCONGRUOUS_ELEMENTS_DEGREE = 6, /* bspline/error.rs:202 */
  ENTERP_INVALID_DEGREE = 7,              /* bspline/error.rs:95 */
  ENTERP_TOO_SMALL_WORKSPACE = 8,         /* builder.rs:163 */
  ENTERP_EMPTY = 9,                       /* builder.rs:85 */
  ENTERP_INVALID_ARGUMENT = 10,           /* malformed descriptor / NULL pointer / bad enum */
  ENTERP_UNSUPPORTED = 11,                /* combination the reference has no impl block for */
  ENTERP_NO_DEVICE = 12,                  /* no CUDA device / device index out of range */
  ENTERP_CUDA_ERROR = 13,                 /* a CUDA runtime call failed; see enterp_last_cuda_error */
  ENTERP_NOT_UPLOADED = 14                /* first use of the curve on this device while `stream` is being captured into a
                                             CUDA graph: the one-time table upload cannot run inside a capture; call
                                             enterp_curve_upload(curve, device) first */
} enterp_status;

enum { ENTERP_LINEAR = 0, ENTERP_BEZIER = 1, ENTERP_BSPLINE = 2 };           /* desc.kind      */
enum { ENTERP_F32 = 0, ENTERP_F64 = 1 };                                       /* desc.scalar    */
enum { ENTERP_OPEN = 0, ENTERP_CLAMPED = 1, ENTERP_LEGACY = 2 };               /* desc.mode      */
/* desc.knot_base. CONST_EQUIDISTANT is ConstEquidistant<R, N> (src/base/list.rs:566-721): knots i/(N-1) on [0,1]
 * and a span computed as x*(N-1) — different roundings from Equidistant's (x-offset)/step. Linear:
 * `Linear::equidistant_unchecked(elements)` (src/linear/mod.rs:193-207), N = n_elements. BSpline:
 * `BSpline::new(elements, ConstEquidistant::<R, N>::new(), space)` (src/bspline/mod.rs:212-235), N = n_knots,
 * mode OPEN only (no builder chain reaches this type; the constructor's checks apply). */
enum { ENTERP_KNOTS_SORTED = 0, ENTERP_KNOTS_EQUIDISTANT = 1, ENTERP_KNOTS_CONST_EQUIDISTANT = 2 };
enum { ENTERP_EQ_BY_DEGREE = 0, ENTERP_EQ_BY_QUANTITY = 1 };                   /* desc.eq_by     */
enum { ENTERP_EQ_DOMAIN = 0, ENTERP_EQ_NORMALIZED = 1, ENTERP_EQ_DISTANCE = 2 }; /* desc.eq_form */
enum { ENTERP_SPACE_CONSTANT = 0, ENTERP_SPACE_DYNAMIC = 1, ENTERP_SPACE_WORKSPACE = 2 }; /* desc.space */
enum { ENTERP_INPUT_NORMALIZED = 0, ENTERP_INPUT_DOMAIN = 1, ENTERP_INPUT_TRANSFORM = 2 };  /* desc.input_form (Bezier) */
/* desc.easing (Linear): the built-in easing functions of src/easing/mod.rs:86-132 and plateau.rs:10-57
 * applied to the lerp factor (linear/mod.rs:127). Arbitrary closures (FuncEase) cannot cross to the device. */
enum { ENTERP_EASE_IDENTITY = 0, ENTERP_EASE_FLIP = 1, ENTERP_EASE_SMOOTHSTART = 2, ENTERP_EASE_SMOOTHEND = 3,
       ENTERP_EASE_SMOOTHSTEP = 4, ENTERP_EASE_SMOOTHERSTEP = 5, ENTERP_EASE_PLATEAU = 6 };

/* One descriptor == one finished builder chain. The fields mirror the typestate the reference's
 * builders resolve to; validation happens in the order the builder methods are called in
 * (elements -> knots|equidistant.degree|quantity -> domain|normalized|distance -> space). */
typedef struct enterp_curve_desc {
  int32_t kind;      /* ENTERP_LINEAR | ENTERP_BEZIER | ENTERP_BSPLINE */
  int32_t scalar;    /* ENTERP_F32 | ENTERP_F64: type R of knots, parameters and element components */
  int32_t dim;       /* 1..4 components per element (f32/f64 scalar or [R;2..4]) */
  int32_t weighted;  /* 0: .elements(E); 1: .elements_with_weights(G) -> Weighted<..> output is projected */
  uint64_t n_elements;
  const void* elements; /* host AoS [n_elements][dim] of R */
  const void* weights;  /* host [n_elements] of R, or NULL when !weighted */

  int32_t mode;      /* BSpline: ENTERP_OPEN (default) | ENTERP_CLAMPED | ENTERP_LEGACY */
  int32_t knot_base; /* ENTERP_KNOTS_SORTED: .knots(K); ENTERP_KNOTS_EQUIDISTANT: .equidistant::<R>() */
  uint64_t n_knots;  /* SORTED only */
  const void* knots; /* SORTED only: host [n_knots] of R exactly as the user gave them */

  int32_t eq_by;     /* BSpline+EQUIDISTANT: .degree(eq_count) or .quantity(eq_count) */
  int32_t eq_form;   /* .domain(eq_a, eq_b) | .normalized() | .distance(start = eq_a, step = eq_b) */
  uint64_t eq_count;
  double eq_a, eq_b; /* carried as f64; converted to R exactly when R = f32 values were given */

  int32_t space;        /* .constant::<N>() | .dynamic() | .workspace(S): BSpline and Bezier */
  int32_t input_form;   /* Bezier: .normalized::<R>() | .domain(in_a, in_b) (TransformInput::normalized_to_domain) |
                           TRANSFORM: a stored TransformInput { addition = in_a, multiplication = in_b } taken verbatim
                           (base/adaptors.rs:113-117; what a deserialised curve carries) */
  uint64_t workspace_n; /* N of constant::<N>() / S.len() of workspace(S); ignored for dynamic */
  double in_a, in_b;

  int32_t easing;       /* Linear: .easing(F), one of ENTERP_EASE_* (default identity) */
  int32_t easing_n;     /* N of smoothstart::<R, N> / smoothend::<R, N> (>= 1) */
  double easing_param;  /* strength of Plateau::new(strength) */
} enterp_curve_desc;

/* Payload of the reference's error structs (found/necessary, elements/knots, index, degree). */
typedef struct enterp_error_info {
  int32_t status;
  int32_t mode;        /* build mode reported by IncongruousElements* (bspline/error.rs:124-130) */
  uint64_t found;      /* TooFew*: found; TooSmallWorkspace: found; InvalidDegree: degree;
                          KnotElementInequality/Incongruous*: number of elements */
  uint64_t necessary;  /* TooSmallWorkspace: necessary; KnotElementInequality/IncongruousElementsKnots:
                          number of knots; IncongruousElementsDegree: degree */
  uint64_t index;      /* NotSorted: index (list.rs:327) */
} enterp_error_info;

/* Resolved shape of a built curve (what the Rust type would carry). */
typedef struct enterp_curve_info {
  int32_t kind, scalar, dim, weighted;
  int32_t knot_base;     /* SORTED | EQUIDISTANT | -1 (Bezier) */
  int32_t adaptor;       /* 0 none, 1 BorderBuffer, 2 BorderDeletion (bspline/adaptors.rs) */
  uint64_t n_elements;
  uint64_t n_knots;      /* virtual knot count = K::len() of the built curve */
  uint64_t degree;       /* BSpline: n_knots - n_elements + 1 (bspline/mod.rs:254); Bezier: n-1; Linear: 1 */
  uint64_t border;       /* BorderBuffer n */
  double domain[2];      /* Curve::domain() */
} enterp_curve_info;

/* Constants the reference's built types carry as fields and the builders computed (their serde form,
 * e.g. src/base/list.rs:354-358, src/base/adaptors.rs:113-117, src/base/space.rs:70-73, src/easing/plateau.rs:10-13):
 * everything a host wrapper needs to write a built curve back out in the reference's own layout. */
typedef struct enterp_curve_resolved {
  uint64_t eq_len;          /* innermost Equidistant { len, step, offset }; 0 when the knots are not equidistant */
  double eq_step, eq_offset;
  int32_t has_transform;    /* Bezier .domain(a, b): TransformInput { addition, multiplication } */
  int32_t easing;           /* ENTERP_EASE_* */
  double in_addition, in_multiplication;
  uint64_t space_len;       /* DynSpace { len } / workspace length; N of ConstSpace */
  double ease_min, ease_max; /* Plateau { min, max } */
} enterp_curve_resolved;

/* ---- construction: *::builder()...build() ------------------------------------------------- */

/* Validates the descriptor with the reference's builder rules and resolves it on the host (no CUDA
 * call). On error *out is NULL, the first failing builder step is reported like the Director
 * variants do (src/bspline/builder.rs:225-241, 377-401, 447-466, 512-534, 633-692, 780-836,
 * 1087-1130; src/linear/builder.rs:165-181, 325-340; src/bezier/builder.rs:134-150, 354-367). */
/* Director semantics: validate only the builder steps up to `stage` (0 = elements, 1 = knots,
 * 2 = workspace/everything) so that a host wrapper can fail AT the offending call like the
 * reference's `*Director` types do (e.g. src/bspline/builder.rs:225-241). */
ENTERP_API enterp_status enterp_curve_validate(const enterp_curve_desc* desc, int stage,
                                               enterp_error_info* err /* may be NULL */);
ENTERP_API enterp_status enterp_curve_create(const enterp_curve_desc* desc, enterp_curve** out,
                                             enterp_error_info* err /* may be NULL */);
ENTERP_API void enterp_curve_destroy(enterp_curve* curve);
ENTERP_API enterp_status enterp_curve_get_info(const enterp_curve* curve, enterp_curve_info* info);
ENTERP_API enterp_status enterp_curve_get_resolved(const enterp_curve* curve, enterp_curve_resolved* out);
/* Curve::domain(): linear/mod.rs:139-141, bspline/mod.rs:180-185, bezier/mod.rs:257-259,
 * base/adaptors.rs:161-166 (TransformInput). */
ENTERP_API enterp_status enterp_curve_domain(const enterp_curve* curve, double out[2]);
/* Curve::clamp() (src/base/signal.rs:271-290 -> Clamp, src/base/adaptors.rs:14-44): a new handle that
 * clamps the parameter to `base`'s domain before evaluating; shares base's tables. domain() is unchanged. */
ENTERP_API enterp_status enterp_curve_clamp(const enterp_curve* base, enterp_curve** out);
/* Curve::slice(bounds) (signal.rs:241-262 -> Slice, adaptors.rs:55-106): the parameter is mapped by
 * t * scale + (start - domain_start), scale = (end - start) / (domain_end - domain_start); an absent bound
 * (has_* == 0) defaults to the domain's. domain() is reported as the base's (adaptors.rs:99-105). */
ENTERP_API enterp_status enterp_curve_slice(const enterp_curve* base, int has_start, double start, int has_end,
                                            double end, enterp_curve** out);
/* Replicates the resolved tables onto `device` ("uploaded once"). Implicit on first eval. */
ENTERP_API enterp_status enterp_curve_upload(enterp_curve* curve, int device);

/* ---- evaluation over device-resident buffers ---------------------------------------------- */

/* Signal::sample(iter): out[i] = curve.eval(t[i]); t_dev [n] of R, out_dev [n][dim] of R. */
ENTERP_API enterp_status enterp_eval_batch(enterp_curve* curve, int device, const void* t_dev,
                                           uint64_t n, void* out_dev, void* stream);
/* Curve::take(n_total) restricted to global sample indices [first, first+count): the stepper
 * t_i = step*fl(i)+start with step=(end-start)/fl(n_total-1) (list.rs:393-399, 416-418) is
 * generated on the device from the GLOBAL index, so shards of one take() agree bit for bit. */
ENTERP_API enterp_status enterp_eval_take(enterp_curve* curve, int device, uint64_t n_total,
                                          uint64_t first, uint64_t count, void* out_dev,
                                          void* stream);
/* Parity/debug: the knot-span decision only. BSpline: idx[i] = strict_upper_bound_clamped(t,
 * degree, len-degree) (bspline/mod.rs:148-154); Linear: idx[i] = max_index of upper_border
 * (list.rs:169-203, 532-551); Bezier: 0. factor_dev (may be NULL) receives Linear's factor. */
ENTERP_API enterp_status enterp_span_batch(enterp_curve* curve, int device, const void* t_dev,
                                           uint64_t n, uint32_t* idx_dev, void* factor_dev,
                                           void* stream);
ENTERP_API enterp_status enterp_span_take(enterp_curve* curve, int device, uint64_t n_total,
                                          uint64_t first, uint64_t count, uint32_t* idx_dev,
                                          void* factor_dev, void* stream);
/* The take() parameter values themselves (Stepper, signal.rs:607-609). */
ENTERP_API enterp_status enterp_stepper(int scalar, int device, double start, double end,
                                        uint64_t n_total, uint64_t first, uint64_t count,
                                        void* t_out_dev, void* stream);
/* Many independent Bezier curves sharing one take(samples_per_curve) table on [0,1]:
 * ctrl_dev [n_curves][n_ctrl][dim] of R -> out_dev [n_curves][samples_per_curve][dim].
 * Equivalent to building n_curves `Bezier::builder().elements(ctrl[c]).normalized().constant()`
 * curves and collecting `.take(samples_per_curve)` of each (bezier/mod.rs:240-246). */
ENTERP_API enterp_status enterp_bezier_many_take(int scalar, int dim, uint64_t n_curves,
                                                 uint32_t n_ctrl, const void* ctrl_dev,
                                                 uint32_t samples_per_curve, int device,
                                                 void* out_dev, void* stream);

/* Bezier::gen_with_tangent (bezier/mod.rs:270-276): out_dev [n][2][dim] = (value, tangent) per parameter, and
 * Bezier::gen_with_deriatives::<K> (279-285): out_dev [n][k][dim] = (value, 1st, .., (k-1)-th derivative),
 * bug-for-bug as upstream (the min(k, n_elements-1)-th entry is always zero). Plain Bezier curves only
 * (unweighted, `.normalized()`): anything else returns ENTERP_UNSUPPORTED. k < n_elements indexes out of
 * bounds upstream (a panic) and returns ENTERP_INVALID_ARGUMENT; n_elements, k <= 64. */
ENTERP_API enterp_status enterp_bezier_tangent_batch(enterp_curve* curve, int device, const void* t_dev, uint64_t n,
                                                     void* out_dev, void* stream);
ENTERP_API enterp_status enterp_bezier_derivatives_batch(enterp_curve* curve, int device, const void* t_dev,
                                                         uint64_t n, uint32_t k, void* out_dev, void* stream);

/* ---- evaluation with HOST buffers (the call a user of the reference makes) ----------------- */

/* `curve.sample(t).collect::<Vec<_>>()`: pageable or pinned host in/out, chunked and pipelined
 * H2D -> kernel -> D2H on two streams; blocks until out_host is complete. */
ENTERP_API enterp_status enterp_sample_host(enterp_curve* curve, int device, const void* t_host,
                                            uint64_t n, void* out_host);
/* `curve.take(n_total).skip(first).take(count).collect()`. */
ENTERP_API enterp_status enterp_take_host(enterp_curve* curve, int device, uint64_t n_total,
                                          uint64_t first, uint64_t count, void* out_host);
/* The same two calls spread over several GPUs of the box by ONE call from one host thread: the sample
 * range is cut into n_devices contiguous shares (share g = [g*count/n, (g+1)*count/n), evaluated with the GLOBAL
 * stepper index so the result equals the single-device call bit for bit), every device keeps a replica of the curve
 * (uploaded on first use) and runs its own double-buffered kernel/D2H pipeline; chunks are issued round-robin. No
 * collective is involved. `devices` must not repeat an ordinal. */
ENTERP_API enterp_status enterp_sample_host_multi(enterp_curve* curve, const int* devices, int n_devices,
                                                  const void* t_host, uint64_t n, void* out_host);
ENTERP_API enterp_status enterp_take_host_multi(enterp_curve* curve, const int* devices, int n_devices,
                                                uint64_t n_total, uint64_t first, uint64_t count, void* out_host);
/* Run-length transfer: an f32 take repeats every parameter in runs of 2^(floor(log2 i) - 23) samples beyond index 2^24
 * (Equidistant::eval computes `step * R::from_usize(i) + offset`, src/base/list.rs:416-418, and from_usize keeps 24
 * bits). For such ranges (from 2^26 on; curve without input adaptors, finite domain) the take_host calls evaluate and
 * copy only the DISTINCT values over PCIe and replicate them into out_host on the host with a few threads
 * (ENTERP_HOST_THREADS, default: the calling thread's CPU affinity, at most 16) — byte copies only, the result is the
 * same array bit for bit. enterp_host_transfer_bytes reports the bytes the host-buffer calls have actually moved over
 * PCIe so far in this process (either pointer may be NULL). */
ENTERP_API void enterp_host_transfer_bytes(unsigned long long* h2d_bytes, unsigned long long* d2h_bytes);
/* The host-buffer calls borrow pooled scratch (2 device chunk buffers, 2 streams, optional pinned staging) that is
 * kept for the next call. This frees every idle pipe of `device` (all devices when device < 0); returns the count. */
ENTERP_API int enterp_host_scratch_release(int device);
/* Pinned host memory for zero-staging transfers of the calls above. */
ENTERP_API enterp_status enterp_host_alloc(void** ptr, uint64_t bytes);
ENTERP_API void enterp_host_free(void* ptr);

/* ---- diagnostics --------------------------------------------------------------------------- */

/* Number of samples so far (on `device`) for which the reference would have panicked. NOTE: the eval entry points
 * return ENTERP_OK even when samples were invalid (they come back as NaN); the counter is cumulative per device
 * replica and shared by the handles derived from one curve (clamp / slice). enterp_invalid_count synchronises the
 * whole device; the _stream form only `stream`, and enterp_invalid_reset zeroes the counter in stream order, so
 * reset -> eval -> count_stream on one stream attributes invalid samples to one call. */
ENTERP_API enterp_status enterp_invalid_count(enterp_curve* curve, int device,
                                              unsigned long long* n);
ENTERP_API enterp_status enterp_invalid_count_stream(enterp_curve* curve, int device, void* stream,
                                                     unsigned long long* n);
ENTERP_API enterp_status enterp_invalid_reset(enterp_curve* curve, int device, void* stream);
ENTERP_API const char* enterp_status_str(enterp_status status);
ENTERP_API const char* enterp_last_cuda_error(void);
ENTERP_API int enterp_device_count(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
ENTERP_API unsigned long long enterp_launch_count(void);
ENTERP_API const char* enterp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ENTERP_H_ */
