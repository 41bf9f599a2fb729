//! Headless equivalent of the reference binary (`src/main.rs`): `initial_state` (lines 180-328)
//! followed by the `LBMSim` frame loop step x speed_factor -> render (lines 128-177), without the
//! piston window — the shape of `display::record` (`src/display.rs:157-185`).
//! UNCOMPILED in this repository; the C++ twin `chemsim_b200/cpp/main_rs_harness.cpp` is what runs.
use chemsim_lbm_b200::lbm::{self, CollisionOperator, Scalar};
use chemsim_lbm_b200::matrix;
use chemsim_lbm_b200::af_compat as af;

struct LBMSim {
    speed_factor: usize,
    size: (usize, usize),
    state: lbm::State<lbm::D2Q9>,
}

fn initial_state(size: (usize, usize)) -> LBMSim {
    let (w, h) = size;
    let disc = lbm::Discretization { delta_x: 1.0, delta_t: 1.0 };              // main.rs:185

    // let collision = lbm::BGK { tau: 15.0 };                                    // main.rs:187
    let viscosity = 10.0;
    let collision = lbm::Regularized::new(lbm::KBC::new(viscosity));            // main.rs:198-199

    let initial_velocity = {                                                    // main.rs:201-218
        let mut vec_x: Vec<Scalar> = vec![0.0; w * h];
        let mut vec_y: Vec<Scalar> = vec![0.0; w * h];
        for x in 0..w {
            for y in 0..h {
                vec_x[(y * w) + x] = 0.02;
                vec_y[(y * w) + x] = 0.0;
            }
        }
        (matrix::Matrix::new(&vec_x, size).unwrap(), matrix::Matrix::new(&vec_y, size).unwrap())
    };
    let initial_density = matrix::Matrix::new_filled(1.0, size);                // main.rs:223

    let pops = &({                                                              // main.rs:257-265, verbatim
        let temp = lbm::compute_equilibrium(initial_density, initial_velocity, &lbm::D2Q9::directions(), disc);
        temp.iter().map(|(_, pop)| pop.clone()).collect::<Vec<lbm::Population>>()
    });
    let lattice = lbm::D2Q9::new(pops);                                         // main.rs:267

    let geometry = {                                                            // main.rs:269-312
        let mut vec = vec![false; w * h];
        for x in 0..w {
            for y in 0..h {
                let mut r = 0.0f64;
                r += (x as f64 - (w as f64 / 2.0)).powi(2);
                r += (y as f64 - (h as f64 / 2.0)).powi(2);
                r = r.sqrt();
                if r < 25.0 { vec[y * w + x] = true; }
                if x == 0 || y == 0 || x == w - 1 || y == h - 1 { vec[y * w + x] = true; }
            }
        }
        let dim4 = af::Dim4::new(&[w as u64, h as u64, 1, 1]);                  // main.rs:308-311, verbatim
        af::transpose(&af::Array::new(&vec[..], dim4), false)
    };

    let collision: Box<dyn CollisionOperator<lbm::D2Q9>> = Box::new(collision);
    let state = lbm::State::initial(Box::new(lattice), geometry, collision, disc);   // main.rs:314-319
    LBMSim { size, state, speed_factor: 2 }
}

fn main() {
    let mut sim = initial_state((400, 400));                                    // main.rs:343, :349
    let start = std::time::Instant::now();
    let frames = 500;
    for frame in 0..frames {
        for _ in 0..sim.speed_factor {                                          // main.rs:129-135
            sim.state.step();
        }
        if frame == 250 { sim.state.paint_brush([200.5, 120.5]); }               // main.rs:71-91, on the device
        let image = sim.state.render_rgba(0, true);                             // main.rs:157-176, on the device
        if frame % 100 == 0 {
            println!("frame {} time {} mass {} first pixel {:?}", frame, sim.state.time, sim.state.total_mass(),
                     &image[0..4]);
        }
    }
    println!("Average frames per second: {}", frames as f64 / start.elapsed().as_secs_f64());
    let _ = sim.size;
}
