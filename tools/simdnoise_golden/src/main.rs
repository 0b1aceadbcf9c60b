//! Golden vectors from simdnoise 3.1.6, produced through exactly the calls the reference makes
//! (src/perlin_noise.rs:5-41): `NoiseBuilder::fbm_3d_offset(..).with_seed(1).with_freq(0.09)
//! .with_gain(2.5).with_lacunarity(0.5).with_octaves(3).wrap()` into `avx2::get_3d_noise` and
//! `sse41::get_3d_noise`.  Output: one JSON document on stdout; every f32 is written as the hex of
//! its bit pattern so the comparison in tests/test_simdnoise_reference.py is exact.
use simdnoise::NoiseBuilder;

const CHUNK: usize = 32;

/// CRC-32 (IEEE, the zlib polynomial) so the whole 128 KiB field is pinned, not only the samples.
fn crc32(bytes: &[u8]) -> u32 {
    let mut crc = 0xffff_ffffu32;
    for &b in bytes {
        crc ^= b as u32;
        for _ in 0..8 {
            crc = if crc & 1 != 0 { (crc >> 1) ^ 0xedb8_8320 } else { crc >> 1 };
        }
    }
    !crc
}

fn block(pos: (i32, i32, i32), isa: &str) -> Vec<f32> {
    // perlin_noise.rs:6-19: simdnoise (x, y, z) = engine (x, z, y), origins = pos * 32 as f32
    let generator = NoiseBuilder::fbm_3d_offset(
        (pos.0.wrapping_mul(32)) as f32,
        CHUNK,
        (pos.2.wrapping_mul(32)) as f32,
        CHUNK,
        (pos.1.wrapping_mul(32)) as f32,
        CHUNK,
    )
    .with_seed(1)
    .with_freq(0.09)
    .with_gain(2.5)
    .with_lacunarity(0.5)
    .with_octaves(3)
    .wrap();
    let (noise, _, _) = unsafe {
        match isa {
            "avx2" => simdnoise::avx2::get_3d_noise(&generator), // perlin_noise.rs:20
            _ => simdnoise::sse41::get_3d_noise(&generator),     // perlin_noise.rs:39
        }
    };
    noise
}

fn record(pos: (i32, i32, i32), isa: &str) -> String {
    let noise = block(pos, isa);
    assert_eq!(noise.len(), CHUNK * CHUNK * CHUNK);
    let mut raw = Vec::with_capacity(noise.len() * 4);
    let mut cubes = Vec::with_capacity(noise.len());
    let mut solid = 0usize;
    let mut near_zero = 0usize;
    for v in &noise {
        raw.extend_from_slice(&v.to_bits().to_le_bytes());
        // chunk.rs:38: noise < 0. -> GRASS (0) else AIR (2)
        let s = *v < 0.0;
        cubes.push(if s { 0u8 } else { 2u8 });
        solid += s as usize;
        near_zero += (v.abs() <= 1e-6) as usize;
    }
    let samples: Vec<String> = noise.iter().step_by(997).map(|v| format!("\"{:08x}\"", v.to_bits())).collect();
    format!(
        "{{\"coord\": [{}, {}, {}], \"isa\": \"{}\", \"noise_crc32\": \"{:08x}\", \"occ_crc32\": \"{:08x}\", \
         \"solid\": {}, \"near_zero_1e-6\": {}, \"noise_1000\": \"{:08x}\", \"noise_sample_997\": [{}]}}",
        pos.0, pos.1, pos.2, isa, crc32(&raw), crc32(&cubes), solid, near_zero, noise[1000].to_bits(),
        samples.join(", ")
    )
}

fn main() {
    // the coordinates of tests/golden/golden.json plus two far ones
    let coords = [
        (0, 0, 0), (1, 0, 0), (-1, 0, -1), (511, 0, -512), (-16, 0, 15), (7, 1, -3), (0, -2, 0),
        (-512, 0, 511), (100000, 3, -99999),
    ];
    let mut recs = Vec::new();
    for isa in ["avx2", "sse41"].iter() {
        let have = match *isa {
            "avx2" => is_x86_feature_detected!("avx2"), // chunk.rs:29
            _ => is_x86_feature_detected!("sse4.1"),
        };
        if !have {
            eprintln!("skipping {}: not supported by this CPU", isa);
            continue;
        }
        for c in coords.iter() {
            recs.push(record(*c, isa));
        }
    }
    println!("{{\"crate\": \"simdnoise 3.1.6 / simdeez 1.0.7\", \"call\": \"perlin_noise.rs:5-41\", \"chunks\": [\n  {}\n]}}",
             recs.join(",\n  "));
}
