//! Prints what biquad 0.4.2 computes at the two places the reference calls it
//! (src/dsp/parametric_eq.rs:94-113 `Coefficients::<f32>::from_params`, :121 `DirectForm2Transposed::<f32>::run`):
//!
//!   coef <type> <fs bits> <fc bits> <q bits> <gain bits> <b0> <b1> <b2> <a1> <a2>      (f32 bit patterns, hex)
//!   coef <type> ... ERR <OutsideNyquist|NegativeQ|...>
//!   run <preset name> <n bands> <n samples>   followed by one line of n hex words: the cascade's output, sample-outer /
//!                                             band-inner as StereoParametricEQ::process_block (:166-179), zero initial state
use biquad::{Biquad, Coefficients, DirectForm2Transposed, ToHertz, Type};
use std::fmt::Write as _;
use std::fs;

fn filter_type(t: u32, gain_db: f32) -> Type<f32> {
    // FilterType declaration order, src/dsp/parametric_eq.rs:25-35, mapped as in :94-103
    match t {
        0 => Type::PeakingEQ(gain_db),
        1 => Type::LowShelf(gain_db),
        2 => Type::HighShelf(gain_db),
        3 => Type::LowPass,
        4 => Type::HighPass,
        5 => Type::BandPass,
        6 => Type::Notch,
        7 => Type::AllPass,
        _ => panic!("filter type {t}"),
    }
}

fn bits(s: &str) -> f32 {
    f32::from_bits(u32::from_str_radix(s, 16).expect("hex f32"))
}

fn main() {
    let root = std::env::args().nth(1).unwrap_or_else(|| "tests/golden".into());
    let grid = fs::read_to_string(format!("{root}/biquad_grid.txt")).expect("biquad_grid.txt");
    let raw = fs::read(format!("{root}/biquad_input.f32")).expect("biquad_input.f32");
    let input: Vec<f32> = raw.chunks_exact(4).map(|c| f32::from_le_bytes([c[0], c[1], c[2], c[3]])).collect();
    let mut out = String::new();
    let mut lines = grid.lines().filter(|l| !l.starts_with('#') && !l.trim().is_empty()).peekable();
    while let Some(line) = lines.next() {
        let w: Vec<&str> = line.split_whitespace().collect();
        match w[0] {
            // design <type> <fs> <fc> <q> <gain>  (hex bit patterns; a human-readable copy follows a '#')
            "design" => {
                let t: u32 = w[1].parse().unwrap();
                let (fs, fc, q, g) = (bits(w[2]), bits(w[3]), bits(w[4]), bits(w[5]));
                match Coefficients::<f32>::from_params(filter_type(t, g), fs.hz(), fc.hz(), q) {
                    Ok(c) => writeln!(out, "coef {t} {} {} {} {} {:08x} {:08x} {:08x} {:08x} {:08x}", w[2], w[3], w[4], w[5],
                                      c.b0.to_bits(), c.b1.to_bits(), c.b2.to_bits(), c.a1.to_bits(), c.a2.to_bits()).unwrap(),
                    Err(e) => writeln!(out, "coef {t} {} {} {} {} ERR {:?}", w[2], w[3], w[4], w[5], e).unwrap(),
                }
            }
            // cascade <name> <fs> <n bands> <n samples>, then one "band <type> <fc> <q> <gain>" line per band
            "cascade" => {
                let name = w[1];
                let fs = bits(w[2]);
                let n_bands: usize = w[3].parse().unwrap();
                let n: usize = w[4].parse().unwrap();
                let mut filters: Vec<DirectForm2Transposed<f32>> = Vec::new();
                for _ in 0..n_bands {
                    let b: Vec<&str> = lines.next().expect("band line").split_whitespace().collect();
                    assert_eq!(b[0], "band");
                    let t: u32 = b[1].parse().unwrap();
                    let (fc, q, g) = (bits(b[2]), bits(b[3]), bits(b[4]));
                    let c = Coefficients::<f32>::from_params(filter_type(t, g), fs.hz(), fc.hz(), q).unwrap();
                    filters.push(DirectForm2Transposed::<f32>::new(c));
                }
                writeln!(out, "run {name} {n_bands} {n}").unwrap();
                for &x in &input[..n] {
                    let mut v = x;
                    for f in filters.iter_mut() {
                        v = f.run(v);
                    }
                    write!(out, "{:08x} ", v.to_bits()).unwrap();
                }
                out.push('\n');
            }
            other => panic!("unknown grid record {other}"),
        }
    }
    fs::write(format!("{root}/biquad_ref.txt"), out).expect("write biquad_ref.txt");
    eprintln!("wrote {root}/biquad_ref.txt");
}
