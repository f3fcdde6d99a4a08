//! Terrain_b200.zig -- Terrain module variant that consumes a BUILT mesh (SURVEY 8-f rank 2).
//!
//! Same module surface as Terrain/Terrain.zig (`create_terrain(self, core, filename) !SceneNode.Handle`, `init`,
//! `deinit`), different body: instead of uploading the heightmap and recomputing every vertex in the vertex
//! shader each frame (Terrain.zig:21-50,114-126), the mesh -- positions, normals, index buffer -- is built once by
//! mr_terrain_build_full and drawn with drawIndexed (IndexedDraw.zig).  The positions are the WGSL formula's bit
//! for bit, so the picture is the same; the normals are new (simple diffuse shading below).
//!
//! UNCOMPILED in this repository's build image (no Zig toolchain; mach / zigimg are un-vendored URL dependencies).
const std = @import("std");
const math = @import("root").math;
const mach = @import("root").mach;
const img = @import("zigimg");
const Renderer = @import("root").Renderer;
const mods = @import("root").getModules();
const b200 = @import("myrenderer_b200");
const IndexedDraw = @import("IndexedDraw.zig");

pub const mach_module = .terrain;
pub const mach_systems = .{ .init, .deinit };
const Terrain = @This();
pub const Mod = mach.Mod(@This());

/// pos + normal, two 16-byte Vec3 slots = 32 bytes (MR_LAYOUT_TERRAINVERTEX); the real offsets go to the library
/// through b200.layoutOf, whatever the compiler chose.
pub const TerrainVertex = struct {
    pos: math.Vec3,
    normal: math.Vec3,
};

// The vertex stage only transforms: everything Terrain.zig:21-50 computed per frame is in the vertex buffer.
const shader_src =
    \\@group(0) @binding(0) var<uniform> world_xform: mat4x4<f32>;
    \\struct FragPass {
    \\    @builtin(position) pos: vec4<f32>,
    \\    @location(0) color: vec4<f32>,
    \\}
    \\@vertex fn vertex(@location(0) pos: vec3<f32>, @location(1) normal: vec3<f32>) -> FragPass {
    \\    var out: FragPass;
    \\    out.pos = world_xform * vec4(pos, 1.0);
    \\    let shade = 0.35 + 0.65 * max(dot(normal, normalize(vec3(0.4, 1.0, 0.3))), 0.0);
    \\    out.color = vec4(vec3(pos.y) * shade, 1.0);   // Terrain.zig:73 colours by height
    \\    return out;
    \\}
;

const Mesh = struct {
    indices: IndexedDraw.IndexBuffer,
    tile_boxes: []f32, // 8 floats per tile (mr_terrain_tile_bounds), for per-frame culling
    n: u32,
};

renderer: *Renderer,
pipeline: Renderer.Pipeline.Handle,
backend: ?*b200.Context = null,
meshes: std.AutoHashMapUnmanaged(mach.ObjectID, Mesh) = .{},

pub const TILE = 64; // quads per tile side

pub fn create_terrain(self: *@This(), core: *mach.Core, filename: []const u8) !Renderer.SceneNode.Handle {
    // PNG -> u16 texels, as Terrain.zig:89-95
    const image_file = try std.fs.cwd().openFile(filename, .{});
    defer image_file.close();
    var stream_source = std.io.StreamSource{ .file = image_file };
    var image = try img.png.load(&stream_source, core.allocator, .{ .temp_allocator = core.allocator });
    defer image.deinit(core.allocator);
    const n: u32 = @intCast(image.width);
    const texels = image.pixels.grayscale16; // raw u16: the 1 - v/65535 of Terrain.zig:120 is fused into the kernel

    // mapped-at-creation vertex + index buffers, filled by the library (host pointers: the copy is inside the call)
    var vertex_count: u64 = 0;
    var index_count: u64 = 0;
    var bmin: [3]f32 = undefined;
    var bmax: [3]f32 = undefined;
    try b200.check(b200.mr_terrain_describe(n, null, &bmin, &bmax, &vertex_count, &index_count));
    const device: *mach.gpu.Device = self.renderer.device;
    const vbuf = device.createBuffer(&mach.gpu.Buffer.Descriptor{
        .mapped_at_creation = .true,
        .size = vertex_count * @sizeOf(TerrainVertex),
        .usage = .{ .copy_dst = true, .map_write = true, .vertex = true },
    });
    var indices = IndexedDraw.IndexBuffer.new(self.renderer, index_count);
    const layout = b200.layoutOf(TerrainVertex, math.Vec2, math.Vec3, math.Vec4);
    const vmap = vbuf.getMappedRange(u8, 0, vbuf.getSize()).?;
    try b200.check(b200.mr_terrain_build_full(self.backend, @ptrCast(texels.ptr), b200.HEIGHT_U16, n, &layout, null, vmap.ptr, indices.map().?.ptr));
    vbuf.unmap();
    indices.index_buffer.?.unmap();

    // per-tile boxes for culling (SceneNode.zig:96-110 applied per tile by mr_terrain_cull each frame, if wanted)
    var tiles_r: u32 = 0;
    var tiles_c: u32 = 0;
    try b200.check(b200.mr_terrain_tile_count(n, TILE, TILE, &tiles_r, &tiles_c));
    const boxes = try core.allocator.alignedAlloc(f32, 16, @as(usize, tiles_r) * tiles_c * 8);
    try b200.check(b200.mr_terrain_tile_bounds(self.backend, @ptrCast(texels.ptr), b200.HEIGHT_U16, n, TILE, TILE, null, boxes.ptr));

    // scene node with the terrain's own box (Terrain.zig:103-110: the same numbers, from mr_terrain_describe)
    const result = try Renderer.Instance.createNode(.{
        .pipeline = self.pipeline,
        .bounding_box_p0 = math.Vec3.init(bmin[0], bmin[1], bmin[2]),
        .bounding_box_p1 = math.Vec3.init(bmax[0], bmax[1], bmax[2]),
    });
    const instance = result.get_backing();
    instance.set_vertex_buffer(.{ .vertex_buffer = vbuf, .vertex_count = @intCast(vertex_count), .instance_count = 1, .first_vertex = 0, .first_instance = 0 });
    try self.meshes.put(core.allocator, @intFromEnum(instance), .{ .indices = indices, .tile_boxes = boxes, .n = n });
    result.set(.onRender, render_terrain); // indexed draw instead of Instance.render_instance
    return result;
}

fn render_terrain(instance: Renderer.Instance.Handle, pass: *Renderer.SceneNode.NodePass) void {
    const self: *Terrain = &mods.terrain;
    const mesh = self.meshes.get(@intFromEnum(instance)) orelse return;
    IndexedDraw.render_instance_indexed(instance, pass, mesh.indices);
}

pub fn init(self: *Terrain, renderer: *Renderer) !void {
    self.renderer = renderer;
    try b200.check(b200.mr_context_create(0, &self.backend));
    self.pipeline = try Renderer.Pipeline.create(.{
        .bindings = &.{.{ .location = 0, .type = .{ .builtin = .transform } }},
        .vertex_source = shader_src,
        .vertex_layout = Renderer.VertexLayout.create(TerrainVertex), // the pipeline now has a vertex layout (Pipeline.zig:137)
    });
}

pub fn deinit(self: *@This(), renderer: *Renderer) void {
    var it = self.meshes.valueIterator();
    while (it.next()) |mesh| {
        mesh.indices.free();
        mods.mach_core.allocator.free(mesh.tile_boxes);
    }
    self.meshes.deinit(mods.mach_core.allocator);
    self.pipeline.destroy(renderer);
    _ = b200.mr_context_destroy(self.backend);
}
