//! Known-answer generator that runs the REFERENCE's own Polygon/Triangulation.zig (unmodified apart from the
//! two edits scripts/zig_golden.sh applies: the edge order comes from an explicit (offset, prime) pair instead
//! of std.crypto.random).  Prints one JSON object per line: {"polygon": name, "offset": o, "prime": p, "ids": [...]}.
//! UNCOMPILED in this repository's build image (no Zig toolchain there); written against Zig 0.14.0-dev.2577,
//! the reference's minimum_zig_version (build.zig.zon:18).
const std = @import("std");
const Triangulation = @import("Triangulation.zig");
const unirand = @import("unirand.zig");
const Point = Triangulation.Point;

// App/App.zig:68-83
const polygon1 = [_]Point{ .{ 62.742857, 106.97143 }, .{ 93.085712, 65.828571 }, .{ 147.08571, 85.628572 }, .{ 122.14285, 144.77143 }, .{ 102.34286, 93.857142 }, .{ 79.199998, 130.37143 }, .{ 81.00000, 105.17143 } };
const polygon2 = [_]Point{ .{ 10.0, 10.0 }, .{ 40.0, 10.0 }, .{ 40.0, 40.0 }, .{ 10.0, 40.0 } };

const Sink = struct {
    ids: std.ArrayList(u32),
    pts: []const Point,
};

// the emit callback only sees coordinates (Triangulation.zig:450); the points of both polygons are distinct
fn emit(sink: *Sink, p: Point) void {
    for (sink.pts, 0..) |q, i| {
        if (q[0] == p[0] and q[1] == p[1]) {
            sink.ids.append(@intCast(i)) catch unreachable;
            return;
        }
    }
    sink.ids.append(0xFFFFFFFF) catch unreachable;
}

fn run(allocator: std.mem.Allocator, out: anytype, name: []const u8, pts: []const Point) !void {
    const n: u32 = @intCast(pts.len);
    const primes = [_]u32{ 1, 2, 3, 5 };
    var tri = Triangulation.new(allocator);
    defer tri.destroy();
    var off: u32 = 0;
    while (off < n) : (off += 1) {
        for (primes) |prime| {
            if (prime != 1 and (prime >= n or n % prime == 0)) continue;
            unirand.forced_offset = off;
            unirand.forced_prime = prime;
            var sink = Sink{ .ids = std.ArrayList(u32).init(allocator), .pts = pts };
            defer sink.ids.deinit();
            try tri.create_polygon(pts, &sink, emit);
            try out.print("{{\"polygon\": \"{s}\", \"offset\": {}, \"prime\": {}, \"ids\": [", .{ name, off, prime });
            for (sink.ids.items, 0..) |id, i| {
                if (i != 0) try out.print(", ", .{});
                try out.print("{}", .{id});
            }
            try out.print("]}}\n", .{});
        }
    }
}

pub fn main() !void {
    var gpa = std.heap.GeneralPurposeAllocator(.{}){};
    defer _ = gpa.deinit();
    const out = std.io.getStdOut().writer();
    try run(gpa.allocator(), out, "polygon1", &polygon1);
    try run(gpa.allocator(), out, "polygon2", &polygon2);
}
