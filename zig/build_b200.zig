//! build_b200.zig -- the lines build.zig needs to link the geometry backend (call from build.zig after
//! `exe` is defined, build.zig:50-58):
//!
//!     const b200 = @import("zig/build_b200.zig");
//!     b200.linkGeometryBackend(b, exe, "third_party/myrenderer_b200");
//!
//! UNCOMPILED in this repository's build image (no Zig toolchain).  Written for Zig 0.14.0-dev
//! (build.zig.zon:18); the same static library is linked and run from C++ by examples/app_scene.cpp.
const std = @import("std");

pub fn linkGeometryBackend(b: *std.Build, exe: *std.Build.Step.Compile, root: []const u8) void {
    // static library of sm_100a kernels, built by `make -C myrenderer_b200/csrc`
    exe.addObjectFile(b.path(b.pathJoin(&.{ root, "myrenderer_b200/lib/libmyrenderer_b200.a" })));
    exe.addIncludePath(b.path(b.pathJoin(&.{ root, "include" })));
    exe.addLibraryPath(.{ .cwd_relative = "/usr/local/cuda/lib64" });
    exe.linkSystemLibrary("cudart");
    exe.linkLibCpp(); // the library is C++ behind its C ABI
    // the Zig declarations of the C ABI, importable as @import("myrenderer_b200")
    exe.root_module.addAnonymousImport("myrenderer_b200", .{ .root_source_file = b.path(b.pathJoin(&.{ root, "zig/myrenderer_b200.zig" })) });
}
