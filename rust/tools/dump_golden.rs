//! dump_golden.rs — turns "parity unpinned" into a pinned check, on a machine that CAN build the
//! reference (nightly Rust + ArrayFire 3.6.1; this repository's build image has neither).
//!
//! It links against the UNMODIFIED reference crate (`chemsim`, /root/reference) and uses nothing
//! but its public API — `lbm::compute_equilibrium`, `D2Q9::{directions, new}`, `State::initial`,
//! `State::step`, `State::populations`, `State::{density, velocity, speed}` — on the scenarios of
//! tests/golden_cases.py that the reference can express (f32, zero-fill edges, square lattices),
//! and writes, per case, the INPUTS it used and the nine populations after N steps as `.npy`:
//!
//!     <out>/<case>_rho.npy  _vx.npy  _vy.npy  (f32, shape (h, w), element (y, x) = slice[y*w + x])
//!     <out>/<case>_solid.npy                  (u8,  shape (h, w))
//!     <out>/<case>_n<N>.npy                   (f32, shape (9, h, w): State::populations in order)
//!     <out>/<case>_n<N>_density.npy, _speed.npy, _vx.npy, _vy.npy     (readouts)
//!     <out>/<case>_n<N>_density_stats.npy     (f64 [mean_all, stdev_all] — pins stdev = population sigma,
//!                                              src/render.rs:41-42)
//!
//! tests/test_reference_golden.py ingests that directory (CHEMSIM_REFERENCE_GOLDEN=<out>, default
//! tests/golden/reference/) and holds the oracle and the CUDA path to it bit for bit.  What the
//! vectors pin: af::convolve2's flip/centre convention (src/lbm.rs:722-724 — a mirrored convention
//! shows up as a point-reflected lattice; chemsim_lbm_set_stream_convention(h, 1) is the switch),
//! af::replace's polarity (:745-747), the op order of the elementwise trees, and stdev_all.
//!
//! Build (inside the reference checkout, with its Cargo.toml):
//!     cp <this repo>/rust/tools/dump_golden.rs  src/bin/dump_golden.rs
//!     cargo +nightly run --release --bin dump_golden -- /tmp/chemsim_golden
//!     cp -r /tmp/chemsim_golden <this repo>/tests/golden/reference
//!     python -m pytest tests/test_reference_golden.py            # CPU: oracle vs reference
//!     python -m pytest tests/test_reference_golden.py -m gpu     # B200: CUDA path vs reference
extern crate arrayfire;
extern crate chemsim;

use arrayfire as af;
use chemsim::lbm;
use chemsim::matrix::Matrix;
use std::fs::File;
use std::io::Write;
use std::path::{Path, PathBuf};

// ---- minimal .npy (format 1.0) writer ---------------------------------------------------------------
fn write_npy(path: &Path, descr: &str, shape: &[usize], bytes: &[u8]) {
    let dims: Vec<String> = shape.iter().map(|d| d.to_string()).collect();
    let shape_s = if shape.len() == 1 { format!("({},)", dims[0]) } else { format!("({})", dims.join(", ")) };
    let mut header = format!("{{'descr': '{}', 'fortran_order': False, 'shape': {}, }}", descr, shape_s);
    let unpadded = 10 + header.len() + 1;
    let pad = (64 - unpadded % 64) % 64;
    header.push_str(&" ".repeat(pad));
    header.push('\n');
    let mut f = File::create(path).expect("create npy");
    f.write_all(b"\x93NUMPY\x01\x00").unwrap();
    f.write_all(&(header.len() as u16).to_le_bytes()).unwrap();
    f.write_all(header.as_bytes()).unwrap();
    f.write_all(bytes).unwrap();
}

fn f32_bytes(v: &[f32]) -> Vec<u8> { v.iter().flat_map(|x| x.to_le_bytes().to_vec()).collect() }
fn f64_bytes(v: &[f64]) -> Vec<u8> { v.iter().flat_map(|x| x.to_le_bytes().to_vec()).collect() }

fn dump_matrix(path: &Path, m: &Matrix) {
    let (w, h) = m.get_shape();
    write_npy(path, "<f4", &[h, w], &f32_bytes(&m.get_underlying()));   // get_underlying: y*w + x (matrix.rs:120-126)
}

// ---- scenarios (tests/golden_cases.py, chemsim_b200/scenarios.py) -------------------------------------
struct Scenario { name: &'static str, n: usize, rho: Vec<f32>, vx: Vec<f32>, vy: Vec<f32>, solid: Vec<bool> }

/// main.rs `initial_state((n, n))` (src/main.rs:180-328) with the given disc radius, with or without walls.
fn main_rs(name: &'static str, n: usize, radius: f64, walls: bool) -> Scenario {
    let (w, h) = (n, n);
    let mut solid = vec![false; w * h];
    for x in 0..w {
        for y in 0..h {
            let mut r = 0.0f64;
            r += (x as f64 - (w as f64 / 2.0)).powi(2);
            r += (y as f64 - (h as f64 / 2.0)).powi(2);
            r = r.sqrt();
            if r < radius { solid[y * w + x] = true; }
            if walls && (x == 0 || y == 0 || x == w - 1 || y == h - 1) { solid[y * w + x] = true; }
        }
    }
    Scenario { name, n, rho: vec![1.0; w * h], vx: vec![0.02; w * h], vy: vec![0.0; w * h], solid }
}

/// A seedless, non-symmetric field with rational arithmetic only (no libm: identical on every host):
/// rho = 1 + ((3x + 5y) mod 17 - 8)/400, vx = ((7x + 2y) mod 13 - 6)/200, vy = ((x + 11y) mod 19 - 9)/300,
/// solid where (5x + 3y) mod 23 == 0 or on the border column x == 0.
fn lattice_hash(name: &'static str, n: usize) -> Scenario {
    let mut s = Scenario { name, n, rho: vec![0.0; n * n], vx: vec![0.0; n * n], vy: vec![0.0; n * n], solid: vec![false; n * n] };
    for y in 0..n {
        for x in 0..n {
            let i = y * n + x;
            s.rho[i] = 1.0 + (((3 * x + 5 * y) % 17) as f32 - 8.0) / 400.0;
            s.vx[i] = (((7 * x + 2 * y) % 13) as f32 - 6.0) / 200.0;
            s.vy[i] = (((x + 11 * y) % 19) as f32 - 9.0) / 300.0;
            s.solid[i] = (5 * x + 3 * y) % 23 == 0 || x == 0;
        }
    }
    s
}

fn geometry_of(s: &Scenario) -> lbm::Geometry {
    // exactly main.rs:308-311: Array::new with dims [w, h] (column-major), then transpose
    let dim4 = af::Dim4::new(&[s.n as u64, s.n as u64, 1, 1]);
    af::transpose(&af::Array::new(&s.solid[..], dim4), false)
}

fn run(out: &Path, s: &Scenario, op_name: &str, collision: Box<lbm::CollisionOperator<lbm::D2Q9>>, steps: &[usize]) {
    let size = (s.n, s.n);
    let disc = lbm::Discretization { delta_x: 1.0, delta_t: 1.0 };
    let case = format!("{}_zerofill_{}_float32", s.name, op_name);
    let m = |v: &Vec<f32>| Matrix::new(&v[..], size).unwrap();
    write_npy(&out.join(format!("{}_rho.npy", case)), "<f4", &[s.n, s.n], &f32_bytes(&s.rho));
    write_npy(&out.join(format!("{}_vx.npy", case)), "<f4", &[s.n, s.n], &f32_bytes(&s.vx));
    write_npy(&out.join(format!("{}_vy.npy", case)), "<f4", &[s.n, s.n], &f32_bytes(&s.vy));
    let solid_u8: Vec<u8> = s.solid.iter().map(|&b| b as u8).collect();
    write_npy(&out.join(format!("{}_solid.npy", case)), "|u1", &[s.n, s.n], &solid_u8);

    let pops: Vec<lbm::Population> = lbm::compute_equilibrium(m(&s.rho), (m(&s.vx), m(&s.vy)), &lbm::D2Q9::directions(), disc)
        .iter().map(|(_, p)| p.clone()).collect();                       // main.rs:257-265
    let mut state = lbm::State::initial(Box::new(lbm::D2Q9::new(&pops)), geometry_of(s), collision, disc);
    let mut done = 0;
    for &n in steps {
        while done < n { state.step(); done += 1; }
        let mut all: Vec<f32> = Vec::with_capacity(9 * s.n * s.n);
        for (_, pop) in state.populations().iter() { all.extend(pop.get_underlying()); }
        write_npy(&out.join(format!("{}_n{}.npy", case, n)), "<f4", &[9, s.n, s.n], &f32_bytes(&all));
        let density = state.density();
        dump_matrix(&out.join(format!("{}_n{}_density.npy", case, n)), &density);
        dump_matrix(&out.join(format!("{}_n{}_speed.npy", case, n)), &state.speed());
        let (vx, vy) = state.velocity();
        dump_matrix(&out.join(format!("{}_n{}_vx.npy", case, n)), &vx);
        dump_matrix(&out.join(format!("{}_n{}_vy.npy", case, n)), &vy);
        let stats = [af::mean_all(density.get_array()).0, af::stdev_all(density.get_array()).0];   // render.rs:41-42
        write_npy(&out.join(format!("{}_n{}_density_stats.npy", case, n)), "<f8", &[2], &f64_bytes(&stats));
    }
    println!("{}: {:?} steps", case, steps);
}

fn main() {
    let out = PathBuf::from(std::env::args().nth(1).unwrap_or_else(|| "chemsim_golden".to_string()));
    std::fs::create_dir_all(&out).unwrap();
    af::init();
    af::set_backend(af::Backend::CPU);       // the parity target is the reference's CPU path (no FMA contraction)
    let disc = lbm::Discretization { delta_x: 1.0, delta_t: 1.0 };
    let scenarios = vec![main_rs("mainrs48", 48, 6.0, true), main_rs("mainrs48open", 48, 6.0, false), lattice_hash("hash40", 40)];
    for s in &scenarios {
        run(&out, s, "bgk15", Box::new(lbm::BGK { tau: 15.0 }), &[1, 2, 10]);
        run(&out, s, "bgk08", Box::new(lbm::BGK { tau: 0.8 }), &[1, 5]);
        run(&out, s, "trt", Box::new(lbm::TRT::new(0.25, 0.1, &disc)), &[3]);
        run(&out, s, "regularized", Box::new(lbm::Regularized::new(lbm::KBC::new(10.0))), &[1, 2, 10]);   // main.rs:198-199
        run(&out, s, "kbc", Box::new(lbm::KBC::new(0.1)), &[3]);
    }
}
