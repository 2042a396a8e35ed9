//! `src/b200.rs` of enterpolation 0.3.0 — the module a maintainer of the reference adds (behind `feature = "b200"`,
//! with `enterpolation-b200` as an optional dependency) so that curves built by the reference's own builders move
//! to the GPU: `let gpu = curve.to_device()?; let samples = gpu.take_vec(1 << 24)?;`.
//!
//! The curve structs keep their fields private (`src/bspline/mod.rs:62-67`, `src/linear/mod.rs:66-70`,
//! `src/bezier/mod.rs:162-166`, `src/base/list.rs:255, 354-358`, `src/bspline/adaptors.rs:7-10, 94-96`), so this
//! patch also changes those fields to `pub(crate)`. Nothing else of the crate changes; evaluation on the CPU is
//! untouched. Chains are copied element by element through `Chain::eval(i)` (`src/base/mod.rs:19-70`), which is how
//! every accepted container (`Vec<T>`, `&[T]`, `[T; N]`, user chains) hands out its elements.
//!
//! Source only: no Rust toolchain exists where this repository is built.

use crate::bezier::Bezier;
use crate::bspline::{BSpline, BorderBuffer, BorderDeletion};
use crate::easing::Identity;
use crate::linear::Linear;
use crate::weights::{Weighted, Weights};
use crate::{Chain, Equidistant, Signal, Sorted, Space};

use enterpolation_b200::{Description, DeviceCurve, Element, Error, Offload, Scalar};

fn collect<C: Chain>(chain: &C) -> Vec<C::Output> {
    (0..chain.len()).map(|i| chain.eval(i)).collect()
}

/// Knot chains the engine understands, flattened into the description the way the builders produced them.
pub trait DeviceKnots<R: Scalar> {
    fn describe<T: Element<R>>(&self, d: Description<R, T>, n_elements: usize) -> Description<R, T>;
}

impl<R: Scalar, C: Chain<Output = R>> DeviceKnots<R> for Sorted<C> {
    fn describe<T: Element<R>>(&self, d: Description<R, T>, _n: usize) -> Description<R, T> {
        d.knots(collect(&self.0)) // open mode: the knots as given (bspline/builder.rs:377-401)
    }
}
impl<R: Scalar> DeviceKnots<R> for Equidistant<R> {
    fn describe<T: Element<R>>(&self, d: Description<R, T>, _n: usize) -> Description<R, T> {
        // the resolved Equidistant { len, step, offset } taken verbatim (list.rs:354-408)
        d.equidistant().quantity(self.len as u64).distance(self.offset, self.step)
    }
}
impl<R: Scalar, C: Chain<Output = R>> DeviceKnots<R> for BorderBuffer<Sorted<C>> {
    fn describe<T: Element<R>>(&self, d: Description<R, T>, _n: usize) -> Description<R, T> {
        d.clamped().knots(collect(&self.inner.0)) // n = elements - knots is recomputed by the library (builder.rs:459)
    }
}
impl<R: Scalar> DeviceKnots<R> for BorderBuffer<Equidistant<R>> {
    fn describe<T: Element<R>>(&self, d: Description<R, T>, _n: usize) -> Description<R, T> {
        d.clamped().equidistant().quantity(self.inner.len as u64).distance(self.inner.offset, self.inner.step)
    }
}
impl<R: Scalar, C: Chain<Output = R>> DeviceKnots<R> for BorderDeletion<Sorted<C>> {
    fn describe<T: Element<R>>(&self, d: Description<R, T>, _n: usize) -> Description<R, T> {
        d.legacy().knots(collect(&self.inner.0))
    }
}

// ---- BSpline ---------------------------------------------------------------------------------------------------
impl<R, T, K, E, S> Offload<R, T> for BSpline<K, E, S>
where
    R: Scalar,
    T: Element<R>,
    K: DeviceKnots<R>,
    E: Chain<Output = T>,
    S: Space<T>,
{
    fn to_device(&self) -> Result<DeviceCurve<R, T>, Error> {
        let d = Description::bspline().elements(collect(&self.elements));
        // every Space is performance-equivalent on the GPU; only the builder's size check sees its length
        self.knots.describe(d, self.elements.len()).workspace(self.space.len() as u64).build()
    }
}

// NURBS: Weighted<BSpline<K, Weights<G>, S>> where G yields (element, weight) pairs (weights/mod.rs:19-41)
impl<R, T, K, G, S> Offload<R, T> for Weighted<BSpline<K, Weights<G>, S>>
where
    R: Scalar,
    T: Element<R>,
    K: DeviceKnots<R>,
    G: Chain<Output = (T, R)>,
    S: Space<crate::weights::Homogeneous<T, R>>,
{
    fn to_device(&self) -> Result<DeviceCurve<R, T>, Error> {
        let inner = &self.inner;
        let d = Description::bspline().elements_with_weights(collect(&inner.elements.signal));
        inner.knots.describe(d, inner.elements.len()).workspace(inner.space.len() as u64).build()
    }
}

// ---- Linear ----------------------------------------------------------------------------------------------------
impl<R, T, C, E> Offload<R, T> for Linear<Sorted<C>, E, Identity>
where
    R: Scalar,
    T: Element<R>,
    C: Chain<Output = R>,
    E: Chain<Output = T>,
{
    fn to_device(&self) -> Result<DeviceCurve<R, T>, Error> {
        Description::linear().elements(collect(&self.elements)).knots(collect(&self.knots.0)).build()
    }
}
impl<R, T, E> Offload<R, T> for Linear<Equidistant<R>, E, Identity>
where
    R: Scalar,
    T: Element<R>,
    E: Chain<Output = T>,
{
    fn to_device(&self) -> Result<DeviceCurve<R, T>, Error> {
        Description::linear()
            .elements(collect(&self.elements))
            .equidistant()
            .distance(self.knots.offset, self.knots.step)
            .build()
    }
}

// ---- Bezier ----------------------------------------------------------------------------------------------------
impl<R, T, E, S> Offload<R, T> for Bezier<R, E, S>
where
    R: Scalar,
    T: Element<R>,
    E: Chain<Output = T>,
    S: Space<T>,
{
    fn to_device(&self) -> Result<DeviceCurve<R, T>, Error> {
        Description::bezier().elements(collect(&self.elements)).normalized().workspace(self.space.len() as u64).build()
    }
}

#[cfg(test)]
mod test {
    use super::*;
    use crate::Curve;

    /// README.md:52-62: the device curve reproduces the CPU curve bit for bit.
    #[test]
    fn readme_cubic_on_device() {
        let cpu = BSpline::builder()
            .clamped()
            .elements([0.0, 0.0, 0.0, 6.0, 0.0, 0.0, 0.0])
            .equidistant::<f64>()
            .degree(3)
            .domain(-2.0, 2.0)
            .constant::<4>()
            .build()
            .unwrap();
        let gpu = cpu.to_device().unwrap();
        assert_eq!(gpu.domain(), cpu.domain());
        let want: Vec<f64> = cpu.by_ref().take(4097).collect();
        let got = gpu.take_vec(4097).unwrap();
        assert!(want.iter().zip(&got).all(|(a, b)| a.to_bits() == b.to_bits()));
    }
}
