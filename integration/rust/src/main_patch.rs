// src/main.rs::do_tracing (main.rs:134-180) with the GPU arm.  Everything else in main.rs is unchanged.
//
//     println!("P3\n{} {}\n255", params.render.image_width, params.render.image_height);
//     let start_time = Instant::now();
// -   let rt = Renderer::new_with_rng(camera, world, background, params.render,
// -                                   RecursiveRayTracer { max_depth: params.max_depth }, rngator);
// -   let image = rt.render(|_, total| { ... });
// +   let cam = gpu::RtCamera { lookfrom: params.lookfrom.e, lookat: params.lookat.e, vup: params.up.e,
// +                             vfov_deg: params.field_of_view, aspect_ratio: params.aspect_ratio,
// +                             aperture: params.aperture, focus_dist: params.focus_dist };
// +   let image = gpu::render_gpu(&sink, root, background_kind, top, bottom, &cam,
// +                               params.render.image_width, params.render.image_height,
// +                               params.render.samples_per_pixel, params.max_depth, seed, gpus);
//     eprintln!("\nRendered in {:.3}s", start_time.elapsed().as_secs_f32());
//     for line in image.iter().rev() {
//         for (r, g, b) in line.iter() { println!("{} {} {}", r, g, b); }
//     }
//
// `sink`/`root` come from `World::build(&mut rng, &mut sink)`: each `worlds.rs` recipe gets the extra parameter and
// one `sink.*` call next to each constructor (`mu-lambda-raytracer_b200/csrc/worlds.cpp` is that, for all ten worlds).
