// b200_state.rs — the reference's `State` (src/state.rs:29-833) with the wgpu device / pipeline
// layer replaced by the B200 backend. It keeps the reference's own CPU code for everything that
// feeds the render step — `Scene`, `build_bvh`, `PackedMeshes`, `EnvironmentMaps`, `CameraUniform`
// — and hands their `encase` bytes to the C ABI unchanged (include/rt_types.h mirrors the layouts).
//
// NOT COMPILED in this repository's build image (no Rust toolchain). It is written against the
// reference at the revision surveyed in SURVEY.md and against src/rt_sys.rs, which IS generated from
// the C headers and checked by the test suite; treat this file as the integration recipe.
//
// How the reference's calls map:
//   State::new     create_buffer_init of the scene bind group (state.rs:394-458)  -> rt_upload_scene
//                  EnvironmentMaps::new textures + alias tables (state.rs:119-142) -> rt_upload_environments
//   State::resize  HdrPipeline textures (state.rs:651-700)                          -> rt_resize
//   State::update  camera / env index / dev index uniforms (state.rs:722-758)       -> rt_set_camera (+ fields)
//   State::render  scene-hash reset, sample_count, dispatch_workgroups (760-833)    -> rt_clear / rt_render
//   hdr.process    tone map into the surface (hdr.rs:162-181)                        -> rt_tonemap_rgb8
use std::collections::hash_map::DefaultHasher;
use std::ffi::CStr;
use std::hash::{Hash, Hasher};
use std::ptr;

use crate::bvh::build_bvh;
use crate::camera::CameraUniform;
use crate::rt_sys::*;
use crate::scene::{Scene, SceneState};

/// shader.wgsl:232
pub const MAX_BOUNCES: u32 = 10;

pub struct B200State {
    ctx: *mut RtContext,
    pub scene: SceneState,
    pub width: u32,
    pub height: u32,
    /// HdrPipeline::sample_count (hdr.rs:70): the `sample_count` uniform of the next frame
    pub sample_count: u32,
    pub last_frame_scene_hash: Option<u64>,
}

fn check(status: i32, ctx: *const RtContext) -> anyhow::Result<()> {
    if status == RT_OK {
        return Ok(());
    }
    let text = unsafe { CStr::from_ptr(rt_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!("rt backend error {status}: {text}"))
}

fn storage_bytes<T: encase::ShaderType + encase::internal::WriteInto>(value: &T) -> Vec<u8> {
    encase::StorageBuffer::<()>::content_of::<_, Vec<u8>>(value).unwrap()
}

impl B200State {
    /// `hdris`: the decoded environment images in slot order — what `State::new` hands to
    /// `EnvironmentMaps::new` (state.rs:119-131).
    pub fn new(scene: Scene, hdris: &[image::Rgb32FImage], width: u32, height: u32, device: i32) -> anyhow::Result<Self> {
        let mut ctx: *mut RtContext = ptr::null_mut();
        check(unsafe { rt_create(device, &mut ctx) }, ptr::null())?;

        // The same byte vectors the reference uploads with create_buffer_init (state.rs:394-458).
        let packed = &scene.meshes; // PackedMeshes (mesh.rs:85-113)
        let (primitives, bvh_nodes) = build_bvh(&scene); // or rt_bvh_build for a large scene, see below
        let materials = storage_bytes(&scene.materials);
        let spheres = storage_bytes(&scene.spheres);
        let planes = storage_bytes(&scene.planes.iter().map(|p| p.to_uniform()).collect::<Vec<_>>());
        let vertices = storage_bytes(&packed.vertices());
        let normals = storage_bytes(&packed.normals());
        let triangles = storage_bytes(&packed.triangles());
        let primitive_bytes = storage_bytes(&primitives);
        let node_bytes = storage_bytes(&bvh_nodes);
        let desc = RtSceneDesc {
            materials: materials.as_ptr() as *const RtMaterial,
            material_count: scene.materials.len() as u32,
            spheres: spheres.as_ptr() as *const RtSphere,
            sphere_count: scene.spheres.len() as u32,
            planes: planes.as_ptr() as *const RtPlane,
            plane_count: scene.planes.len() as u32,
            vertices: vertices.as_ptr() as *const RtVec3,
            vertex_count: packed.vertices().len() as u32,
            normals: normals.as_ptr() as *const RtVec3,
            normal_count: packed.normals().len() as u32,
            triangles: triangles.as_ptr() as *const RtTriangle,
            triangle_count: packed.triangles().len() as u32,
            primitives: primitive_bytes.as_ptr() as *const RtPrimitiveInfo,
            primitive_count: primitives.len() as u32,
            bvh_nodes: node_bytes.as_ptr() as *const RtBvhNode,
            bvh_node_count: bvh_nodes.len() as u32,
        };
        check(unsafe { rt_upload_scene(ctx, &desc) }, ctx)?;

        // What EnvironmentMaps::new builds per slot (environments.rs:19-64): rgba32f texels with alpha 0
        // (Texture::from_rgb32float, texture.rs:104-118) and the alias table. `AliasTable` is private to
        // environments.rs, so the table comes from the host library's rt_alias_build, which restates
        // AliasTable::build_by_luminance (environments.rs:97-186) entry for entry.
        let mut texels: Vec<Vec<f32>> = Vec::new();
        let mut tables: Vec<Vec<RtAliasEntry>> = Vec::new();
        for hdri in hdris {
            let (w, h) = (hdri.width(), hdri.height());
            let rgb: &[f32] = hdri.as_flat_samples().samples;
            let mut rgba = vec![0f32; (w * h * 4) as usize];
            let mut table = Vec::<RtAliasEntry>::with_capacity((w * h) as usize);
            let mut leftover = 0u32;
            unsafe {
                rt_rgb_to_rgba(w * h, rgb.as_ptr(), rgba.as_mut_ptr());
                check(rt_alias_build(w, h, rgb.as_ptr(), table.as_mut_ptr(), &mut leftover), ptr::null())?;
                table.set_len((w * h) as usize);
            }
            log::info!("AliasTable: {} left over pixels out of {}", leftover, w * h);
            texels.push(rgba);
            tables.push(table);
        }
        let envs: Vec<RtEnvironmentDesc> = hdris
            .iter()
            .enumerate()
            .map(|(i, hdri)| RtEnvironmentDesc {
                width: hdri.width(),
                height: hdri.height(),
                rgba: texels[i].as_ptr(),
                alias: tables[i].as_ptr(),
            })
            .collect();
        check(unsafe { rt_upload_environments(ctx, envs.as_ptr(), envs.len() as u32) }, ctx)?;
        check(unsafe { rt_resize(ctx, width, height) }, ctx)?;

        let mut state = Self {
            ctx,
            scene: SceneState { camera: scene.camera.clone(), environment_index: 0, dev_index: 0 },
            width,
            height,
            sample_count: 0,
            last_frame_scene_hash: None,
        };
        state.update()?;
        Ok(state)
    }

    /// State::resize (state.rs:651-700): new accumulation target, accumulation restarts.
    pub fn resize(&mut self, width: u32, height: u32) -> anyhow::Result<()> {
        if width == 0 || height == 0 {
            return Ok(());
        }
        check(unsafe { rt_resize(self.ctx, width, height) }, self.ctx)?;
        self.width = width;
        self.height = height;
        self.sample_count = 0;
        self.last_frame_scene_hash = None;
        Ok(())
    }

    /// State::update (state.rs:722-758): the camera uniform. Resolution, environment index and
    /// `dev_index` are arguments of the calls below instead of uniforms.
    pub fn update(&mut self) -> anyhow::Result<()> {
        let bytes = storage_bytes(&CameraUniform::new(&self.scene.camera));
        check(unsafe { rt_set_camera(self.ctx, bytes.as_ptr() as *const RtCameraUniform) }, self.ctx)
    }

    /// State::render (state.rs:760-833): one frame = one sample. A changed scene hash clears the
    /// accumulation and restarts at sample 0, otherwise the sample index advances — the value
    /// that seeds the per-pixel RNG (shader.wgsl:1309-1312). The call returns at once; the frame is
    /// complete when `read_accum` / `present` return.
    pub fn render(&mut self) -> anyhow::Result<()> {
        let mut hasher = DefaultHasher::new();
        self.scene.hash(&mut hasher);
        let hash = hasher.finish();
        if Some(hash) != self.last_frame_scene_hash {
            self.last_frame_scene_hash = Some(hash);
            self.sample_count = 0;
            check(unsafe { rt_clear(self.ctx) }, self.ctx)?;
        } else {
            self.sample_count += 1;
        }
        check(
            unsafe { rt_render(self.ctx, self.sample_count, 1, MAX_BOUNCES, self.scene.environment_index) },
            self.ctx,
        )
    }

    /// Headless form: `count` consecutive frames in one call (what the benchmark times).
    pub fn render_samples(&mut self, count: u32, max_bounces: u32) -> anyhow::Result<()> {
        check(
            unsafe { rt_render(self.ctx, self.sample_count, count, max_bounces, self.scene.environment_index) },
            self.ctx,
        )?;
        self.sample_count += count;
        Ok(())
    }

    /// hdr.process (hdr.rs:162-181) without a surface: accumulation / n -> f16 -> ACES -> sRGB, 8 bit.
    pub fn present(&mut self) -> anyhow::Result<Vec<u8>> {
        let mut rgb8 = vec![0u8; (self.width * self.height * 3) as usize];
        check(unsafe { rt_tonemap_rgb8(self.ctx, self.sample_count + 1, rgb8.as_mut_ptr()) }, self.ctx)?;
        Ok(rgb8)
    }

    /// The cumulative_light_texture (rgba32f, alpha = 1 where rendered).
    pub fn read_accum(&mut self) -> anyhow::Result<Vec<f32>> {
        let mut rgba = vec![0f32; (self.width * self.height * 4) as usize];
        check(unsafe { rt_read_accum(self.ctx, rgba.as_mut_ptr()) }, self.ctx)?;
        Ok(rgba)
    }

    /// The `dev_index` 2 / 3 views of `main` (shader.wgsl:1314-1338).
    pub fn debug_view(&mut self, dev_index: u32) -> anyhow::Result<Vec<f32>> {
        let mut rgba = vec![0f32; (self.width * self.height * 4) as usize];
        check(
            unsafe {
                rt_debug_view(self.ctx, dev_index, self.scene.environment_index, self.sample_count, rgba.as_mut_ptr())
            },
            self.ctx,
        )?;
        Ok(rgba)
    }
}

impl Drop for B200State {
    fn drop(&mut self) {
        unsafe { rt_destroy(self.ctx) };
    }
}

/// build_bvh (bvh.rs:13-28) on the GPU for a scene too large for the recursive CPU builder: the same
/// node array and primitive order, byte for byte. `plane_bounds` holds `Plane::bounds()` of every
/// plane as [min.x, min.y, min.z, max.x, max.y, max.z] (scene.rs:203-207).
pub fn build_bvh_on_device(
    ctx: *mut RtContext,
    spheres: &[u8],
    sphere_count: u32,
    plane_bounds: &[f32],
    vertices: &[u8],
    vertex_count: u32,
    triangles: &[u8],
    triangle_count: u32,
) -> anyhow::Result<(Vec<RtPrimitiveInfo>, Vec<RtBvhNode>, u32)> {
    let input = RtBvhBuildInput {
        spheres: spheres.as_ptr() as *const RtSphere,
        sphere_count,
        plane_bounds: plane_bounds.as_ptr(),
        plane_count: (plane_bounds.len() / 6) as u32,
        vertices: vertices.as_ptr() as *const RtVec3,
        vertex_count,
        triangles: triangles.as_ptr() as *const RtTriangle,
        triangle_count,
    };
    let (mut node_count, mut depth) = (0u32, 0u32);
    check(unsafe { rt_bvh_build(ctx, &input, &mut node_count, &mut depth, ptr::null_mut()) }, ctx)?;
    let primitive_count = (sphere_count + input.plane_count + triangle_count) as usize;
    let mut nodes = Vec::<RtBvhNode>::with_capacity(node_count as usize);
    let mut primitives = Vec::<RtPrimitiveInfo>::with_capacity(primitive_count);
    check(unsafe { rt_bvh_build_fetch(ctx, nodes.as_mut_ptr(), primitives.as_mut_ptr()) }, ctx)?;
    unsafe {
        nodes.set_len(node_count as usize);
        primitives.set_len(primitive_count);
    }
    Ok((primitives, nodes, depth))
}
