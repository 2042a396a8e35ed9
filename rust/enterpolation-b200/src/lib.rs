//! Device-resident curves for [enterpolation]: the host side of the B200-native batched curve-evaluation engine.
//!
//! A curve is described once (the same information the reference's typestate builders resolve to), uploaded once,
//! and evaluated on the GPU through `libenterp_b200.so` (C ABI: `include/enterp.h`, raw bindings:
//! `enterpolation-b200-sys`). [`DeviceCurve`] implements the reference's own [`Signal`] and [`Curve`] traits
//! (`src/base/signal.rs:15-19, 187-228`), so it drops into code that is generic over them; the batched calls
//! ([`DeviceCurve::sample_vec`], [`DeviceCurve::take_vec`], [`DeviceCurve::take_vec_multi`]) are the hot path:
//! `curve.sample(ts).collect::<Vec<_>>()` and `curve.take(n).collect::<Vec<_>>()` of the reference.
//!
//! Results are bit-identical to the reference's CPU `eval` (same IEEE operations in the same order; where the
//! reference panics — NaN/inf into `to_usize().unwrap()` — the sample comes back as NaN and
//! [`DeviceCurve::invalid_count`] counts it). There is no CPU fallback: without the library or a B200 every
//! evaluation returns [`Error::Device`].
//!
//! No Rust toolchain exists in the image this repository is developed in: this crate is source that mirrors,
//! call for call, the C++ (`include/enterpolation.hpp`) and Python (`enterpolation_b200/`) host layers, which are
//! compiled and tested there.
//!
//! [enterpolation]: https://docs.rs/enterpolation/0.3.0

use core::marker::PhantomData;
use std::ffi::CStr;
use std::os::raw::c_void;

use enterpolation::{Curve, Signal};
pub use enterpolation_b200_sys as sys;

/// The scalar type `R` of knots, parameters and element components.
pub trait Scalar: num_traits::real::Real + num_traits::FromPrimitive + Copy + 'static {
    /// `ENTERP_F32` / `ENTERP_F64`.
    const CODE: i32;
    /// The ABI carries scalar constants as `f64`: exact for `f64`, the nearest `f32` for `f32` (the values it carries
    /// back were widened from `f32` by the library, so nothing is lost). Named apart from num-traits' `from_f64`/`to_f64`.
    fn narrow(v: f64) -> Self;
    fn widen(self) -> f64;
}
impl Scalar for f32 {
    const CODE: i32 = sys::ENTERP_F32;
    fn narrow(v: f64) -> Self {
        v as f32
    }
    fn widen(self) -> f64 {
        self as f64
    }
}
impl Scalar for f64 {
    const CODE: i32 = sys::ENTERP_F64;
    fn narrow(v: f64) -> Self {
        v
    }
    fn widen(self) -> f64 {
        self
    }
}

/// Element types the device evaluates: `R` itself and `[R; 2..=4]` (dense arrays of `DIM` scalars).
///
/// # Safety
/// `Self` must have exactly the layout of `[R; DIM]` (size, alignment of `R`, no padding, every bit pattern valid).
/// Implement it for your own vector/colour newtypes (`#[repr(transparent)]` over `[R; D]`) whose `Merge` is the
/// component-wise `a * (1 - f) + b * f` of topology-traits' blanket impl; anything else cannot be offloaded.
pub unsafe trait Element<R: Scalar>: Copy {
    const DIM: usize;
}
unsafe impl Element<f32> for f32 {
    const DIM: usize = 1;
}
unsafe impl Element<f64> for f64 {
    const DIM: usize = 1;
}
unsafe impl<R: Scalar, const D: usize> Element<R> for [R; D] {
    const DIM: usize = D;
}

/// Builder errors of the reference (`LinearError`, `BezierError`, `BSplineError` and their payload structs) plus
/// the ABI-level conditions.
#[derive(Debug, Clone, PartialEq)]
pub enum Error {
    /// `TooFewElements { found }` (`src/builder.rs:107`)
    TooFewElements { found: u64 },
    /// `TooFewKnots { found }` (`src/builder.rs:135`)
    TooFewKnots { found: u64 },
    /// `KnotElementInequality { elements, knots }` (`src/linear/error.rs:56`)
    KnotElementInequality { elements: u64, knots: u64 },
    /// `NotSorted { index }` (`src/base/list.rs:325`)
    NotSorted { index: u64 },
    /// `IncongruousElementsKnots { elements, knots, mode }` (`src/bspline/error.rs:132`); mode = `ENTERP_OPEN|CLAMPED|LEGACY`
    IncongruousElementsKnots { elements: u64, knots: u64, mode: i32 },
    /// `IncongruousElementsDegree { elements, degree, mode }` (`src/bspline/error.rs:202`)
    IncongruousElementsDegree { elements: u64, degree: u64, mode: i32 },
    /// `InvalidDegree { degree }` (`src/bspline/error.rs:95`)
    InvalidDegree { degree: u64 },
    /// `TooSmallWorkspace { found, necessary }` (`src/builder.rs:163`)
    TooSmallWorkspace { found: u64, necessary: u64 },
    /// `Empty` (`src/builder.rs:85`)
    Empty,
    /// Malformed description (NULL pointer, bad enum, `dim` outside 1..=4, `take(0)`, range outside the take).
    InvalidArgument,
    /// A combination the reference has no `impl` block for, or a size beyond the engine's limits.
    Unsupported,
    /// No usable B200 / a CUDA runtime failure (message of `enterp_last_cuda_error`). There is no CPU fallback.
    Device(String),
    /// First use of a curve on a device while the given stream is being captured into a CUDA graph.
    NotUploaded,
}

impl core::fmt::Display for Error {
    fn fmt(&self, f: &mut core::fmt::Formatter<'_>) -> core::fmt::Result {
        write!(f, "{self:?}")
    }
}
impl std::error::Error for Error {}

fn check(status: sys::enterp_status, info: &sys::enterp_error_info) -> Result<(), Error> {
    Err(match status {
        sys::ENTERP_OK => return Ok(()),
        sys::ENTERP_TOO_FEW_ELEMENTS => Error::TooFewElements { found: info.found },
        sys::ENTERP_TOO_FEW_KNOTS => Error::TooFewKnots { found: info.found },
        sys::ENTERP_KNOT_ELEMENT_INEQUALITY => Error::KnotElementInequality { elements: info.found, knots: info.necessary },
        sys::ENTERP_NOT_SORTED => Error::NotSorted { index: info.index },
        sys::ENTERP_INCONGRUOUS_ELEMENTS_KNOTS => {
            Error::IncongruousElementsKnots { elements: info.found, knots: info.necessary, mode: info.mode }
        }
        sys::ENTERP_INCONGRUOUS_ELEMENTS_DEGREE => {
            Error::IncongruousElementsDegree { elements: info.found, degree: info.necessary, mode: info.mode }
        }
        sys::ENTERP_INVALID_DEGREE => Error::InvalidDegree { degree: info.found },
        sys::ENTERP_TOO_SMALL_WORKSPACE => Error::TooSmallWorkspace { found: info.found, necessary: info.necessary },
        sys::ENTERP_EMPTY => Error::Empty,
        sys::ENTERP_INVALID_ARGUMENT => Error::InvalidArgument,
        sys::ENTERP_UNSUPPORTED => Error::Unsupported,
        sys::ENTERP_NOT_UPLOADED => Error::NotUploaded,
        _ => {
            // SAFETY: the library returns a NUL-terminated thread-local string that lives until the next failing call
            let msg = unsafe { CStr::from_ptr(sys::enterp_last_cuda_error()) }.to_string_lossy().into_owned();
            Error::Device(msg)
        }
    })
}

fn no_info() -> sys::enterp_error_info {
    sys::enterp_error_info { status: 0, mode: 0, found: 0, necessary: 0, index: 0 }
}

/// Built-in easing functions of `Linear` (`src/easing/mod.rs:86-132`, `src/easing/plateau.rs`); arbitrary closures
/// (`FuncEase`) cannot cross to the device.
#[derive(Debug, Clone, Copy, PartialEq)]
pub enum Easing {
    Identity,
    Flip,
    SmoothStart(u32),
    SmoothEnd(u32),
    SmoothStep,
    SmootherStep,
    /// `Plateau::new(strength)`
    Plateau(f64),
}

/// One finished builder chain, owned. The methods carry the names (and the call order) of the reference's
/// builders: `Linear::builder()` (`src/linear/mod.rs:105`), `Bezier::builder()` (`src/bezier/mod.rs:207`),
/// `BSpline::builder()` (`src/bspline/mod.rs:115`). Validation happens in the library with the reference's rules and
/// error payloads; `build()` reports the first failing step like the `*Builder` types do.
pub struct Description<R: Scalar, T: Element<R>> {
    d: sys::enterp_curve_desc,
    elements: Vec<T>,
    weights: Vec<R>,
    knots: Vec<R>,
    device: i32,
}

impl<R: Scalar, T: Element<R>> Description<R, T> {
    fn new(kind: i32) -> Self {
        let d = sys::enterp_curve_desc {
            kind,
            scalar: R::CODE,
            dim: T::DIM as i32,
            weighted: 0,
            n_elements: 0,
            elements: core::ptr::null(),
            weights: core::ptr::null(),
            mode: sys::ENTERP_OPEN,
            knot_base: sys::ENTERP_KNOTS_SORTED,
            n_knots: 0,
            knots: core::ptr::null(),
            eq_by: sys::ENTERP_EQ_BY_DEGREE,
            eq_form: sys::ENTERP_EQ_NORMALIZED,
            eq_count: 0,
            eq_a: 0.0,
            eq_b: 1.0,
            space: sys::ENTERP_SPACE_DYNAMIC,
            input_form: sys::ENTERP_INPUT_NORMALIZED,
            workspace_n: 0,
            in_a: 0.0,
            in_b: 1.0,
            easing: sys::ENTERP_EASE_IDENTITY,
            easing_n: 0,
            easing_param: 0.0,
        };
        Description { d, elements: Vec::new(), weights: Vec::new(), knots: Vec::new(), device: 0 }
    }
    /// `Linear::builder()`
    pub fn linear() -> Self {
        Self::new(sys::ENTERP_LINEAR)
    }
    /// `Bezier::builder()`
    pub fn bezier() -> Self {
        Self::new(sys::ENTERP_BEZIER)
    }
    /// `BSpline::builder()` (open mode, dynamic workspace unless changed)
    pub fn bspline() -> Self {
        Self::new(sys::ENTERP_BSPLINE)
    }
    /// `.clamped()` (`src/bspline/builder.rs:177-209`)
    pub fn clamped(mut self) -> Self {
        self.d.mode = sys::ENTERP_CLAMPED;
        self
    }
    /// `.legacy()`
    pub fn legacy(mut self) -> Self {
        self.d.mode = sys::ENTERP_LEGACY;
        self
    }
    /// `.elements(E)`
    pub fn elements(mut self, elements: impl Into<Vec<T>>) -> Self {
        self.elements = elements.into();
        self.weights.clear();
        self.d.weighted = 0;
        self
    }
    /// `.elements_with_weights(G)`: the built curve is `Weighted<..>` (NURBS), its output the projected element.
    pub fn elements_with_weights(mut self, pairs: impl IntoIterator<Item = (T, R)>) -> Self {
        self.elements.clear();
        self.weights.clear();
        for (e, w) in pairs {
            self.elements.push(e);
            self.weights.push(w);
        }
        self.d.weighted = 1;
        self
    }
    /// `.knots(K)`: sorted knots exactly as the user gives them.
    pub fn knots(mut self, knots: impl Into<Vec<R>>) -> Self {
        self.knots = knots.into();
        self.d.knot_base = sys::ENTERP_KNOTS_SORTED;
        self
    }
    /// `.equidistant::<R>()` — follow with `degree`/`quantity` (B-splines) and `domain`/`normalized`/`distance`.
    pub fn equidistant(mut self) -> Self {
        self.knots.clear();
        self.d.knot_base = sys::ENTERP_KNOTS_EQUIDISTANT;
        self
    }
    /// `ConstEquidistant<R, N>` knots: `Linear::equidistant_unchecked` (N = number of elements; pass 0) or
    /// `BSpline::new(elements, ConstEquidistant::<R, N>::new(), space)` (pass N).
    pub fn const_equidistant(mut self, n: u64) -> Self {
        self.knots.clear();
        self.d.knot_base = sys::ENTERP_KNOTS_CONST_EQUIDISTANT;
        self.d.n_knots = n;
        self
    }
    /// `.degree(d)` (`src/bspline/builder.rs:633-655, 780-798`)
    pub fn degree(mut self, degree: u64) -> Self {
        self.d.eq_by = sys::ENTERP_EQ_BY_DEGREE;
        self.d.eq_count = degree;
        self
    }
    /// `.quantity(q)` (`673-692, 820-836`)
    pub fn quantity(mut self, quantity: u64) -> Self {
        self.d.eq_by = sys::ENTERP_EQ_BY_QUANTITY;
        self.d.eq_count = quantity;
        self
    }
    /// `.domain(start, end)`: of the equidistant knots (Linear, BSpline) or of a Bezier's input (`TransformInput`).
    pub fn domain(mut self, start: R, end: R) -> Self {
        if self.d.kind == sys::ENTERP_BEZIER {
            self.d.input_form = sys::ENTERP_INPUT_DOMAIN;
            self.d.in_a = start.widen();
            self.d.in_b = end.widen();
        } else {
            self.d.eq_form = sys::ENTERP_EQ_DOMAIN;
            self.d.eq_a = start.widen();
            self.d.eq_b = end.widen();
        }
        self
    }
    /// `.normalized()`
    pub fn normalized(mut self) -> Self {
        if self.d.kind == sys::ENTERP_BEZIER {
            self.d.input_form = sys::ENTERP_INPUT_NORMALIZED;
        } else {
            self.d.eq_form = sys::ENTERP_EQ_NORMALIZED;
        }
        self
    }
    /// `.distance(start, step)`
    pub fn distance(mut self, start: R, step: R) -> Self {
        self.d.eq_form = sys::ENTERP_EQ_DISTANCE;
        self.d.eq_a = start.widen();
        self.d.eq_b = step.widen();
        self
    }
    /// `.constant::<N>()`
    pub fn constant(mut self, n: u64) -> Self {
        self.d.space = sys::ENTERP_SPACE_CONSTANT;
        self.d.workspace_n = n;
        self
    }
    /// `.dynamic()`
    pub fn dynamic(mut self) -> Self {
        self.d.space = sys::ENTERP_SPACE_DYNAMIC;
        self
    }
    /// `.workspace(S)` with `S.len() == n`
    pub fn workspace(mut self, n: u64) -> Self {
        self.d.space = sys::ENTERP_SPACE_WORKSPACE;
        self.d.workspace_n = n;
        self
    }
    /// `.easing(F)` for the built-in easings (`src/linear/builder.rs:496-503`)
    pub fn easing(mut self, easing: Easing) -> Self {
        let (kind, n, p) = match easing {
            Easing::Identity => (sys::ENTERP_EASE_IDENTITY, 0, 0.0),
            Easing::Flip => (sys::ENTERP_EASE_FLIP, 0, 0.0),
            Easing::SmoothStart(n) => (sys::ENTERP_EASE_SMOOTHSTART, n as i32, 0.0),
            Easing::SmoothEnd(n) => (sys::ENTERP_EASE_SMOOTHEND, n as i32, 0.0),
            Easing::SmoothStep => (sys::ENTERP_EASE_SMOOTHSTEP, 0, 0.0),
            Easing::SmootherStep => (sys::ENTERP_EASE_SMOOTHERSTEP, 0, 0.0),
            Easing::Plateau(s) => (sys::ENTERP_EASE_PLATEAU, 0, s),
        };
        self.d.easing = kind;
        self.d.easing_n = n;
        self.d.easing_param = p;
        self
    }
    /// Not in the reference: the GPU the built curve evaluates on by default.
    pub fn on_device(mut self, device: i32) -> Self {
        self.device = device;
        self
    }

    fn synced(&mut self) -> *const sys::enterp_curve_desc {
        self.d.n_elements = self.elements.len() as u64;
        self.d.elements = if self.elements.is_empty() { core::ptr::null() } else { self.elements.as_ptr() as *const c_void };
        self.d.weights = if self.d.weighted != 0 && !self.weights.is_empty() {
            self.weights.as_ptr() as *const c_void
        } else {
            core::ptr::null()
        };
        if self.d.knot_base == sys::ENTERP_KNOTS_SORTED {
            self.d.n_knots = self.knots.len() as u64;
            self.d.knots = if self.knots.is_empty() { core::ptr::null() } else { self.knots.as_ptr() as *const c_void };
        } else {
            self.d.knots = core::ptr::null();
        }
        &self.d
    }

    /// `*Director` semantics: validate the chain up to `stage` (0 elements, 1 knots, 2 workspace).
    pub fn validate(&mut self, stage: i32) -> Result<(), Error> {
        let mut info = no_info();
        let d = self.synced();
        // SAFETY: `d` points into `self`, the arrays it names are owned by `self` and outlive the call
        let st = unsafe { sys::enterp_curve_validate(d, stage, &mut info) };
        check(st, &info)
    }

    /// `.build()`: validates like the reference's builders and resolves the curve on the host (no CUDA call; the
    /// tables are uploaded on first use or by [`DeviceCurve::upload`]).
    pub fn build(mut self) -> Result<DeviceCurve<R, T>, Error> {
        let mut info = no_info();
        let mut h: *mut sys::enterp_curve = core::ptr::null_mut();
        let d = self.synced();
        // SAFETY: as in `validate`; the library copies every array during create
        let st = unsafe { sys::enterp_curve_create(d, &mut h, &mut info) };
        check(st, &info)?;
        Ok(DeviceCurve { h, device: self.device, _r: PhantomData })
    }
}

/// A built curve living on the GPU(s): immutable, uploaded once per device, evaluated by hand-written sm_100a
/// kernels. Mirrors `impl Signal<R> + Curve<R>` of `Linear`, `Bezier`, `BSpline` and `Weighted<..>`.
pub struct DeviceCurve<R: Scalar, T: Element<R>> {
    h: *mut sys::enterp_curve,
    device: i32,
    _r: PhantomData<(R, T)>,
}

// SAFETY: the handle is immutable after create and the library allows concurrent calls from several host threads
// (`eval(&self)` re-entrancy of the reference, `src/base/space.rs:63-66`).
unsafe impl<R: Scalar, T: Element<R>> Send for DeviceCurve<R, T> {}
unsafe impl<R: Scalar, T: Element<R>> Sync for DeviceCurve<R, T> {}

impl<R: Scalar, T: Element<R>> Drop for DeviceCurve<R, T> {
    fn drop(&mut self) {
        // SAFETY: `h` came from enterp_curve_create / clamp / slice and is destroyed exactly once
        unsafe { sys::enterp_curve_destroy(self.h) }
    }
}

impl<R: Scalar, T: Element<R>> DeviceCurve<R, T> {
    /// The raw handle, for the device-buffer entry points of `sys` (`enterp_eval_batch`, `enterp_eval_take`, ...).
    pub fn as_raw(&self) -> *mut sys::enterp_curve {
        self.h
    }
    /// Replicates the tables onto `device` now ("uploaded once"); required before capturing a stream into a CUDA graph.
    pub fn upload(&self, device: i32) -> Result<(), Error> {
        // SAFETY: valid handle
        check(unsafe { sys::enterp_curve_upload(self.h, device) }, &no_info())
    }
    fn vec_of(n: usize) -> Vec<T> {
        let mut v = Vec::<T>::with_capacity(n);
        // SAFETY: `T: Element` admits every bit pattern and the caller overwrites all `n` entries before reading
        unsafe { v.set_len(n) };
        v
    }
    /// `curve.sample(ts).collect::<Vec<_>>()` (`src/base/signal.rs:166-173`).
    pub fn sample_vec(&self, ts: &[R]) -> Result<Vec<T>, Error> {
        let mut out = Self::vec_of(ts.len());
        if !ts.is_empty() {
            // SAFETY: `ts` and `out` hold `ts.len()` parameters / elements of the curve's scalar type
            let st = unsafe {
                sys::enterp_sample_host(self.h, self.device, ts.as_ptr() as *const c_void, ts.len() as u64, out.as_mut_ptr() as *mut c_void)
            };
            check(st, &no_info())?;
        }
        Ok(out)
    }
    /// `curve.take(samples).collect::<Vec<_>>()` (`src/base/signal.rs:221-228`).
    pub fn take_vec(&self, samples: usize) -> Result<Vec<T>, Error> {
        self.take_range(samples, 0, samples)
    }
    /// `curve.take(samples).skip(first).take(count).collect()`: one shard of a take; the stepper uses the GLOBAL index
    /// and the global sample count, so shards agree bit for bit with the whole.
    pub fn take_range(&self, samples: usize, first: usize, count: usize) -> Result<Vec<T>, Error> {
        let mut out = Self::vec_of(count);
        if count > 0 {
            // SAFETY: `out` holds `count` elements
            let st = unsafe {
                sys::enterp_take_host(self.h, self.device, samples as u64, first as u64, count as u64, out.as_mut_ptr() as *mut c_void)
            };
            check(st, &no_info())?;
        }
        Ok(out)
    }
    /// The same take spread over several GPUs of the box by one call (`enterp_take_host_multi`).
    pub fn take_vec_multi(&self, samples: usize, devices: &[i32]) -> Result<Vec<T>, Error> {
        let mut out = Self::vec_of(samples);
        if samples > 0 {
            // SAFETY: `devices` holds `devices.len()` ordinals, `out` holds `samples` elements
            let st = unsafe {
                sys::enterp_take_host_multi(self.h, devices.as_ptr(), devices.len() as i32, samples as u64, 0, samples as u64,
                                            out.as_mut_ptr() as *mut c_void)
            };
            check(st, &no_info())?;
        }
        Ok(out)
    }
    /// `Curve::clamp()` (`src/base/signal.rs:271-290`): shares this curve's device tables.
    pub fn clamp(&self) -> Result<Self, Error> {
        let mut h: *mut sys::enterp_curve = core::ptr::null_mut();
        // SAFETY: valid handle, `h` receives a new handle
        check(unsafe { sys::enterp_curve_clamp(self.h, &mut h) }, &no_info())?;
        Ok(DeviceCurve { h, device: self.device, _r: PhantomData })
    }
    /// `Curve::slice(start..end)` (`src/base/signal.rs:241-262`); `None` = unbounded on that side.
    pub fn slice(&self, start: Option<R>, end: Option<R>) -> Result<Self, Error> {
        let mut h: *mut sys::enterp_curve = core::ptr::null_mut();
        // SAFETY: valid handle
        let st = unsafe {
            sys::enterp_curve_slice(self.h, start.is_some() as i32, start.map_or(0.0, Scalar::widen), end.is_some() as i32,
                                    end.map_or(0.0, Scalar::widen), &mut h)
        };
        check(st, &no_info())?;
        Ok(DeviceCurve { h, device: self.device, _r: PhantomData })
    }
    /// Samples so far for which the reference would have panicked (they were written as NaN).
    pub fn invalid_count(&self) -> Result<u64, Error> {
        let mut n: core::ffi::c_ulonglong = 0;
        // SAFETY: valid handle and out pointer
        check(unsafe { sys::enterp_invalid_count(self.h, self.device, &mut n) }, &no_info())?;
        Ok(n as u64)
    }
}

/// `Signal::eval` for one parameter: a 1-sample batch through the GPU path. Panics (like the reference's `eval` may)
/// when no device is usable; use [`DeviceCurve::sample_vec`] for fallible, batched evaluation.
impl<R: Scalar, T: Element<R>> Signal<R> for DeviceCurve<R, T> {
    type Output = T;
    fn eval(&self, input: R) -> T {
        self.sample_vec(&[input]).expect("enterpolation-b200: evaluation failed")[0]
    }
}

impl<R: Scalar, T: Element<R>> Curve<R> for DeviceCurve<R, T> {
    /// `Curve::domain()` (`linear/mod.rs:139-141`, `bspline/mod.rs:180-185`, `bezier/mod.rs:257-259`).
    fn domain(&self) -> [R; 2] {
        let mut d = [0f64; 2];
        // SAFETY: valid handle, `d` has room for two doubles; the call cannot fail for a built curve
        unsafe { sys::enterp_curve_domain(self.h, d.as_mut_ptr()) };
        [R::narrow(d[0]), R::narrow(d[1])]
    }
}

/// Flattening a curve the reference's builders produced into a [`Description`]. The reference keeps the fields of
/// `BSpline`, `Linear` and `Bezier` private, so the impls live inside the reference crate behind a `b200` feature:
/// `rust/reference-patch/b200.rs` of this repository is that module.
pub trait Offload<R: Scalar, T: Element<R>> {
    fn to_device(&self) -> Result<DeviceCurve<R, T>, Error>;
}
