// ref_dump.rs -- pins this repo's oracle (and through it the CUDA path) against the REAL oxipng.
//
// The image this repo was built in has no Rust toolchain, so parity is bit-exact against a C restatement of the reference
// (oracle/) only.  A maintainer with cargo closes that gap with one command: this program runs the reference's own
// `PngImage::filter_image` (src/png/mod.rs:355-431) and reduction entry points (re-exported through
// `oxipng::internal_tests`, src/lib.rs:58-64) on the reference's test images and writes, per image, the raw
// `PngImage.data`, every filtered stream, and the data of every reduction that applies.  tools/compare_ref_dump.py then
// checks those bytes against tests/golden/expected.json (the SHA-256 pins of the oracle the GPU tests are held to).
//
// Usage, from a checkout of oxipng 10.1.1:
//     cp <this repo>/tools/ref_dump.rs examples/ref_dump.rs
//     cargo run --release --example ref_dump -- tests/files /tmp/ref_dump
//     python <this repo>/tools/compare_ref_dump.py /tmp/ref_dump
use std::{env, fs, path::PathBuf};

use oxipng::{internal_tests::*, FilterStrategy, Options};

fn put(dir: &PathBuf, name: &str, what: &str, bytes: &[u8]) {
    fs::write(dir.join(format!("{name}.{what}.bin")), bytes).unwrap();
}

fn main() {
    let args: Vec<String> = env::args().collect();
    let files = PathBuf::from(&args[1]);
    let out = PathBuf::from(&args[2]);
    fs::create_dir_all(&out).unwrap();
    let strategies = [
        ("None", FilterStrategy::NONE), ("Sub", FilterStrategy::SUB), ("Up", FilterStrategy::UP),
        ("Average", FilterStrategy::AVERAGE), ("Paeth", FilterStrategy::PAETH), ("MinSum", FilterStrategy::MinSum),
        ("Entropy", FilterStrategy::Entropy), ("Bigrams", FilterStrategy::Bigrams), ("BigEnt", FilterStrategy::BigEnt),
    ];
    let mut opts = Options::default();
    opts.fix_errors = true;
    for entry in fs::read_dir(&files).unwrap() {
        let path = entry.unwrap().path();
        if path.extension().map_or(true, |e| e != "png") { continue; }
        let name = path.file_stem().unwrap().to_string_lossy().to_string();
        let Ok(png) = PngData::new(&path, &opts) else { continue };
        let img = &png.raw;
        put(&out, &name, "data", &img.data);
        for (sname, s) in &strategies {
            for alpha in [false, true] {
                let (filtered, _used) = img.filter_image(s.clone(), alpha);
                put(&out, &name, &format!("filter.{sname}.{}", alpha as u8), &filtered);
            }
        }
        let red: Vec<(&str, Option<PngImage>)> = vec![
            ("cleaned_alpha_channel", cleaned_alpha_channel(img)),
            ("reduced_alpha_channel.0", reduced_alpha_channel(img, false)),
            ("reduced_alpha_channel.1", reduced_alpha_channel(img, true)),
            ("reduced_bit_depth_16_to_8.0", reduced_bit_depth_16_to_8(img, false)),
            ("reduced_bit_depth_16_to_8.1", reduced_bit_depth_16_to_8(img, true)),
            ("reduced_bit_depth_8_or_less", reduced_bit_depth_8_or_less(img)),
            ("expanded_bit_depth_to_8", expanded_bit_depth_to_8(img)),
            ("reduced_to_indexed.0", reduced_to_indexed(img, false)),
            ("reduced_to_indexed.1", reduced_to_indexed(img, true)),
            ("reduced_rgb_to_grayscale", reduced_rgb_to_grayscale(img)),
            ("reduced_palette.0", reduced_palette(img, false)),
            ("reduced_palette.1", reduced_palette(img, true)),
            ("sorted_palette", sorted_palette(img)),
        ];
        for (rname, r) in red {
            if let Some(r) = r { put(&out, &name, &format!("reduction.{rname}"), &r.data); }
        }
    }
}
