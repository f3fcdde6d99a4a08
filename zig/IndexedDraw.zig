//! IndexedDraw.zig -- the consumer half of the built terrain mesh (SURVEY 8-f rank 2): an index buffer next to
//! Renderer.VertexBuffer and the indexed variant of Instance.render_instance.
//!
//! The reference draws everything with `pass.draw(vertex_count, ...)` (Instance.zig:34-51); the terrain has no
//! vertex buffer at all (vertex pulling, Terrain.zig:126).  A mesh built by mr_terrain_build is indexed, so it
//! needs `setIndexBuffer` + `drawIndexed`.  This file adds exactly that and nothing else; everything it touches
//! is the reference's own API (mach.gpu, Renderer.Instance.Handle, SceneNode.NodePass).
//!
//! UNCOMPILED in this repository's build image (no Zig toolchain; mach is an un-vendored URL dependency).
const mach = @import("root").mach;
const math = @import("root").math;
const Renderer = @import("root").Renderer;

/// Draw descriptor of an indexed mesh: the analogue of Renderer.VertexBuffer (VertexBuffer.zig:5-9) with an
/// index range instead of a vertex range.  `index_count` may be smaller than the buffer holds: after
/// mr_terrain_cull it is counts[1], the indices of the tiles that survived the visibility test.
pub const IndexBuffer = struct {
    index_buffer: ?*mach.gpu.Buffer = null,
    index_count: u32 = 0,
    first_index: u32 = 0,
    base_vertex: i32 = 0,

    /// u32 indices, mapped at creation like VertexBuffer.new (VertexBuffer.zig:14-18) so the library can fill it.
    pub fn new(renderer: *Renderer, index_count: u64) IndexBuffer {
        const device: *mach.gpu.Device = renderer.device;
        const buf = device.createBuffer(&mach.gpu.Buffer.Descriptor{
            .mapped_at_creation = .true,
            .size = index_count * @sizeOf(u32),
            .usage = .{ .copy_dst = true, .map_write = true, .index = true },
        });
        return IndexBuffer{ .index_buffer = buf, .index_count = @intCast(index_count) };
    }

    pub fn map(self: *IndexBuffer) ?[]u32 {
        const buf = self.index_buffer orelse return null;
        return buf.getMappedRange(u32, 0, buf.getSize() / @sizeOf(u32));
    }

    pub fn free(self: *IndexBuffer) void {
        if (self.index_buffer) |buf| buf.release();
        self.index_buffer = null;
    }
};

/// Indexed variant of Instance.render_instance (Instance.zig:34-51): same transform upload, pipeline, vertex
/// buffer and bind group; then setIndexBuffer + drawIndexed instead of draw.  Install it as the scene node's
/// onRender callback (SceneNode.zig:24,118-120) through a small trampoline that looks the IndexBuffer up.
pub fn render_instance_indexed(instance: Renderer.Instance.Handle, pass: *Renderer.SceneNode.NodePass, indices: IndexBuffer) void {
    const pipeline = instance.get_pipeline();
    if (pipeline.get_builtin_location(.transform)) |transform_location| {
        instance.update_buffer(transform_location, 0, math.Mat, &.{pass.xform});
    }
    pass.pass.setPipeline(pipeline.get(.pipeline_handle));
    const vb: Renderer.VertexBuffer = instance.get(.vertex_buffer);
    if (vb.vertex_buffer) |vertex_buffer| {
        pass.pass.setVertexBuffer(0, vertex_buffer, 0, vertex_buffer.getSize());
    }
    pass.pass.setBindGroup(0, instance.get(.bind_group).?, instance.get(.dynamic_offsets));
    const ib = indices.index_buffer orelse return;
    pass.pass.setIndexBuffer(ib, .uint32, 0, ib.getSize());
    pass.pass.drawIndexed(indices.index_count, vb.instance_count, indices.first_index, indices.base_vertex, vb.first_instance);
}
