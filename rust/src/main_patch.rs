//! The tail of `fn main()` from main.rs:939 on.  Everything above it - the scene functions (main.rs:192-795),
//! the `match` arms with their per-scene settings (main.rs:815-937) - stays as it is.
//!
//! Removed: the rayon pixel loop with `ray_color` / `ray_color_without_light_objects` (main.rs:957-1001).
use std::ptr;

use crate::ffi::*;
use crate::flatten::SceneBuilder;

pub fn render_and_print(
    world: &dyn crate::hittable::Hittable,
    lights: &Option<Vec<Box<dyn crate::hittable::Hittable>>>,
    background: crate::color::Color,
    camera: &crate::camera::Camera,
    image_width: usize,
    image_height: usize,
    samples_per_pixel: usize,
    max_depth: usize,
) {
    // flatten: one POD row per constructor call of the scene function
    let mut b = SceneBuilder::default();
    b.world = world.flatten(&mut b);
    if let Some(ls) = lights {
        b.has_lights = true; // Some(vec![]) would panic in choose().unwrap() (hittable.rs:153): the library reports it
        for l in ls {
            let id = l.flatten(&mut b);
            b.lights.push(id);
        }
    }

    // RT1W_GPUS=n: devices 0..n-1 render together (rt1w_context_create_multi: the library shards the sample range over
    // them and adds the radiance sums with one ncclReduce); nothing below changes
    let gpus: i32 = std::env::var("RT1W_GPUS").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
    let mut ctx = ptr::null_mut();
    if gpus > 1 {
        let ids: Vec<i32> = (0..gpus).collect();
        check(unsafe { rt1w_context_create_multi(ids.as_ptr(), gpus, &mut ctx) });
    } else {
        check(unsafe { rt1w_context_create(0, &mut ctx) });
    }
    let mut scene = ptr::null_mut();
    check(unsafe { rt1w_scene_create(ctx, &b.desc(), &mut scene) });

    let params = rt1w_render_params {
        width: image_width as i32,
        height: image_height as i32,
        sample_begin: 0,
        sample_end: samples_per_pixel as i32,
        max_depth: max_depth as i32,
        flags: 0,
        seed: 0,
        background: [background.0.x, background.0.y, background.0.z],
        stat_clamp: 0.0,
        pool_paths: 0,
        reserved: 0,
    };
    let mut rgb8 = vec![0u8; image_width * image_height * 3];
    eprint!("\rScanlines remaining: {} ", image_height); // main.rs:995-998
    check(unsafe { rt1w_render_rgb8(scene, &camera.as_ffi(), &params, rgb8.as_mut_ptr(), ptr::null_mut()) });
    eprint!("\rScanlines remaining: 0 ");

    println!("P3\n{} {}\n255", image_width, image_height); // main.rs:953
    for px in rgb8.chunks(3) {
        // rows arrive top first, as main.rs:959 emits them; the values are `Display for SampledColor` (color.rs:56-65)
        println!("{} {} {}", px[0], px[1], px[2]);
    }
    eprintln!("\nDone"); // main.rs:1009

    unsafe {
        rt1w_scene_destroy(scene);
        rt1w_context_destroy(ctx);
    }
}
