//! Polygon_b200.zig -- the two functions of Polygon/Polygon.zig whose bodies change (init, create_polygon) plus a
//! batched entry point; everything else in that file (GPUVertex, the shader, Handle, deinit) stays as it is.
//! Paste these over Polygon.zig:81-117.  The reusable Triangulation (Polygon.zig:18,116) becomes the library
//! context; render_point's work (palette colour, byte layout, the bbox rule as written, Polygon.zig:65-79) is done
//! by the kernel.
//!
//! UNCOMPILED in this repository's build image (no Zig toolchain; mach is an un-vendored URL dependency).
const std = @import("std");
const mach = @import("root").mach;
const math = @import("root").math;
const Renderer = @import("root").Renderer;
const mods = @import("root").getModules();
const b200 = @import("myrenderer_b200");
const Polygon = @import("Polygon.zig");
const Point = Polygon.Point;

/// fields added to the Polygon module struct
pub const Backend = struct {
    ctx: ?*b200.Context = null,
    layout: b200.Layout = undefined,
    seed: u64 = 0x5EED0003, // replaces std.crypto.random (unirand.zig:31): edge orders are reproducible
    next_index: u64 = 0,
};

pub fn init_backend(be: *Backend, comptime GPUVertex: type) !void {
    try b200.check(b200.mr_context_create(0, &be.ctx));
    be.layout = b200.layoutOf(GPUVertex, math.Vec2, math.Vec3, math.Vec4); // the compiler's real @offsetOf values
}

/// Polygon.create_polygon (Polygon.zig:81-107): one polygon, host memory in and out -> the library's small-batch
/// path (one copy in, one kernel, one copy out).
pub fn create_polygon(self: *Polygon, be: *Backend, comptime GPUVertex: type, vertices: []const Point) !Polygon.Handle {
    var vertex_buffer = Renderer.VertexBuffer.new(self.renderer, 0, @intCast(vertices.len - 2), GPUVertex);
    const mapped = vertex_buffer.map(GPUVertex).?; // the mapped range render_point used to append to
    const first_point = [_]u64{ 0, vertices.len };
    const first_tri = [_]u64{ 0, vertices.len - 2 };
    var bbox: [4]f32 = undefined;
    var status: [1]u32 = .{0};
    const job = b200.PolygonJob{
        .xy = @ptrCast(vertices.ptr),
        .first_point = &first_point,
        .point_base = 0,
        .npoly = 1,
        .offset_prime = null,
        .seed = be.seed,
        .poly_index0 = be.next_index,
        .layout = be.layout,
        .vtx_out = @ptrCast(mapped.ptr),
        .first_tri = &first_tri,
        .tri_base = 0,
        .bbox_out = &bbox,
        .status_out = &status,
        .ntri_out = null,
    };
    try b200.check(b200.mr_triangulate_batch(be.ctx, &job));
    be.next_index += 1;
    // status[0] != 0: one of the situations in which the reference panics or has undefined behaviour
    // (MR_POLY_* in the header); the vertex range is zero-filled beyond what was emitted.
    const node = try Renderer.Instance.createNode(.{
        .pipeline = self.pipeline,
        .bounding_box_p0 = math.Vec3.init(bbox[0], bbox[1], 0.0), // Polygon.zig:96-97
        .bounding_box_p1 = math.Vec3.init(bbox[2], bbox[3], 0.0),
    });
    node.get_backing().set_vertex_buffer(vertex_buffer);
    return @enumFromInt(try self.polygons.new(.{ .vertex_buffer = vertex_buffer, .node = node }));
}

/// Many polygons in one call: one packed vertex buffer, polygon i drawn as the sub-range
/// VertexBuffer{first_vertex = 3*first_tri[i], vertex_count = 3*(n_i - 2)} -- the `offset` / `primitive_count`
/// parameters VertexBuffer.new already has (VertexBuffer.zig:11,21-22).
pub fn create_polygons(self: *Polygon, be: *Backend, comptime GPUVertex: type, allocator: std.mem.Allocator, points: []const Point, first_point: []const u64) ![]Renderer.VertexBuffer {
    const npoly: u32 = @intCast(first_point.len - 1);
    const first_tri = try allocator.alloc(u64, npoly + 1);
    defer allocator.free(first_tri);
    first_tri[0] = 0;
    for (0..npoly) |i| {
        const n = first_point[i + 1] - first_point[i];
        first_tri[i + 1] = first_tri[i] + (if (n >= 2) n - 2 else 0);
    }
    var packed_buffer = Renderer.VertexBuffer.new(self.renderer, 0, @intCast(first_tri[npoly]), GPUVertex);
    const mapped = packed_buffer.map(GPUVertex).?;
    const job = b200.PolygonJob{
        .xy = @ptrCast(points.ptr),
        .first_point = first_point.ptr,
        .point_base = first_point[0],
        .npoly = npoly,
        .offset_prime = null,
        .seed = be.seed,
        .poly_index0 = be.next_index,
        .layout = be.layout,
        .vtx_out = @ptrCast(mapped.ptr),
        .first_tri = first_tri.ptr,
        .tri_base = 0,
        .bbox_out = null,
        .status_out = null,
        .ntri_out = null,
    };
    try b200.check(b200.mr_triangulate_batch(be.ctx, &job));
    be.next_index += npoly;
    const ranges = try allocator.alloc(Renderer.VertexBuffer, npoly);
    for (ranges, 0..) |*r, i| {
        var d: b200.DrawRange = undefined;
        try b200.check(b200.mr_polygon_draw_range(first_tri[i], first_tri[i + 1], 0, &d));
        r.* = .{ .vertex_buffer = packed_buffer.vertex_buffer, .vertex_count = d.vertex_count, .instance_count = d.instance_count, .first_vertex = d.first_vertex, .first_instance = d.first_instance };
    }
    return ranges;
}
