"""Second, independent restatement of Polygon/Triangulation.zig in pure Python.

TEST INFRASTRUCTURE ONLY.  Purpose: cross-check oracle/o_triangulation.c.  It was written
directly from the Zig source, separately from the C file, with Python `None` for Zig `null`
and numpy float32 scalars for every f32 operation.  It is slow (small cases only).

Line numbers refer to Polygon/Triangulation.zig.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
POINT, SEGMENT, TRAPEZOID = "point", "segment", "trapezoid"


class NullUnwrap(Exception):
    pass


def _req(v):
    if v is None:
        raise NullUnwrap()
    return v


class Node:
    __slots__ = ("type", "crumb", "child1", "child2", "point1", "point2")

    def __init__(self, typ):
        self.type = typ
        self.crumb = self.child1 = self.child2 = self.point1 = self.point2 = None

    def copy(self):
        n = Node(self.type)
        n.crumb, n.child1, n.child2, n.point1, n.point2 = (
            self.crumb, self.child1, self.child2, self.point1, self.point2)
        return n


class Triangulation:
    def __init__(self, atan2=None):
        self.nodes = []
        self.points = None
        self.node_stack = []
        self.root_node = None
        self.atan2 = atan2 or (lambda y, x: F(math.atan2(float(y), float(x))))
        self.stuck = False

    # :102-115
    def add_node(self, typ):
        self.nodes.append(Node(typ))
        return len(self.nodes) - 1

    def clone_node(self, node):
        self.nodes.append(self.nodes[node].copy())
        return len(self.nodes) - 1

    # :117-126
    def is_left_of(self, pid, s1, s2):
        p, a, b = self.points[pid], self.points[s1], self.points[s2]
        mul1 = F(F(b[0] - a[0]) * F(p[1] - a[1]))
        mul2 = F(F(b[1] - a[1]) * F(p[0] - a[0]))
        d = F(mul1 - mul2)
        return bool(d > 0)

    # :128-136
    def point_is_above(self, lhs, rhs):
        ly, ry = self.points[lhs][1], self.points[rhs][1]
        if ly < ry:
            return True
        elif ly == ry:
            return bool(self.points[lhs][0] < self.points[rhs][0])
        return False

    # :139-196
    def add_point(self, point_id):
        base_node = self.root_node
        while True:
            nd = self.nodes[base_node]
            if nd.type == TRAPEZOID:
                break
            if nd.type == POINT:
                if nd.point1 == point_id:
                    return
                if self.point_is_above(point_id, _req(nd.point1)):
                    next_node = _req(nd.child1)
                else:
                    next_node = _req(nd.child2)
            else:
                if self.is_left_of(point_id, _req(nd.point1), _req(nd.point2)):
                    next_node = _req(nd.child1)
                else:
                    next_node = _req(nd.child2)
            base_node = next_node
        lower = self.clone_node(base_node)
        upper = self.clone_node(base_node)
        nd = self.nodes[base_node]
        nd.type = POINT
        nd.point1 = point_id
        nd.point2 = None
        nd.crumb = None
        nd.child1 = upper
        nd.child2 = lower
        self.nodes[upper].point2 = point_id
        self.nodes[lower].point1 = point_id

    # :215-396
    def add_segment(self, point1, point2):
        if self.point_is_above(point1, point2):
            upper, lower = point1, point2
        else:
            upper, lower = point2, point1
        base_node = self.root_node
        breadcrumb = None
        self.node_stack = []
        while True:
            while True:
                nd = self.nodes[base_node]
                if nd.type == POINT:
                    pc = _req(nd.point1)
                    if upper == pc:
                        base_node = _req(nd.child2)
                    elif lower == pc:
                        base_node = _req(nd.child1)
                    else:
                        bottom_point_is_above = self.point_is_above(lower, pc)
                        top_point_is_below = self.point_is_above(pc, upper)
                        if top_point_is_below:
                            base_node = _req(nd.child2)
                        elif bottom_point_is_above:
                            base_node = _req(nd.child1)
                        else:
                            nd.crumb = breadcrumb
                            breadcrumb = base_node
                            base_node = _req(nd.child1)
                elif nd.type == SEGMENT:
                    o1, o2 = _req(nd.point1), _req(nd.point2)
                    if upper == o2 or upper == o1:
                        is_left = self.is_left_of(lower, o1, o2)
                    elif lower == o1 or lower == o2:
                        is_left = self.is_left_of(upper, o1, o2)
                    else:
                        top_is_above = self.point_is_above(upper, o1)
                        bottom_is_below = self.point_is_above(lower, o2)
                        if top_is_above and bottom_is_below:
                            is_left = not self.is_left_of(o1, upper, lower)
                        elif top_is_above and not bottom_is_below:
                            is_left = self.is_left_of(lower, o1, o2)
                        else:
                            is_left = self.is_left_of(upper, o1, o2)
                    base_node = _req(nd.child1) if is_left else _req(nd.child2)
                else:
                    break
            self.node_stack.append(base_node)
            if breadcrumb is not None:
                crumb = breadcrumb
                breadcrumb = self.nodes[crumb].crumb
                self.nodes[crumb].crumb = None
                base_node = _req(self.nodes[crumb].child2)
            else:
                break

        left_trapezoid = self.add_node(TRAPEZOID)
        self.nodes[left_trapezoid].point1 = upper
        right_trapezoid = self.add_node(TRAPEZOID)
        self.nodes[right_trapezoid].point1 = upper
        N = self.nodes
        while len(self.node_stack) > 0:
            base_node_index = 0
            base_id = self.node_stack[0]
            low_point = lower
            for i, node in enumerate(self.node_stack):
                new_point = _req(N[node].point2)
                if self.point_is_above(new_point, low_point):
                    low_point = new_point
                    base_node_index = i
                    base_id = node
            N[base_id].type = SEGMENT
            N[left_trapezoid].child1 = N[base_id].child1
            N[base_id].child1 = left_trapezoid
            N[base_id].crumb = left_trapezoid if point1 == upper else right_trapezoid
            N[right_trapezoid].child2 = N[base_id].child2
            N[base_id].child2 = right_trapezoid
            N[base_id].point1 = upper
            N[base_id].point2 = lower
            if lower == low_point:
                N[left_trapezoid].child2 = base_id
                N[left_trapezoid].point2 = low_point
                N[right_trapezoid].child1 = base_id
                N[right_trapezoid].point2 = low_point
                break
            else:
                if self.is_left_of(low_point, upper, lower):
                    N[left_trapezoid].child2 = base_id
                    N[left_trapezoid].point2 = low_point
                    left_trapezoid = self.add_node(TRAPEZOID)
                    N[left_trapezoid].point1 = low_point
                else:
                    N[right_trapezoid].child1 = base_id
                    N[right_trapezoid].point2 = low_point
                    right_trapezoid = self.add_node(TRAPEZOID)
                    N[right_trapezoid].point1 = low_point
            # swapRemove
            last = self.node_stack.pop()
            if base_node_index < len(self.node_stack):
                self.node_stack[base_node_index] = last

    # :398-425
    def push_triangle_if_acute(self, point, axis1, axis2, emit):
        P = self.points
        nx1 = F(P[point][0] - P[axis1][0])
        ny1 = F(P[point][1] - P[axis1][1])
        nx2 = F(P[point][0] - P[axis2][0])
        ny2 = F(P[point][1] - P[axis2][1])
        is_acute = bool(abs(F(self.atan2(ny1, nx1) - self.atan2(ny2, nx2))) < F(math.pi))
        if is_acute:
            emit(point)
            if (axis1 > point and axis2 > point) or (axis1 < point and axis2 < point):
                if axis1 > axis2:
                    emit(axis2)
                    emit(axis1)
                else:
                    emit(axis1)
                    emit(axis2)
            elif axis2 > point:
                emit(axis2)
                emit(axis1)
            elif axis1 > point:
                emit(axis1)
                emit(axis2)
        return is_acute

    # :446-589
    def create_polygon(self, points, edge_order, emit, trace=None):
        self.node_stack = []
        self.nodes = []
        self.points = np.asarray(points, dtype=np.float32)
        n = len(points)
        self.root_node = self.add_node(TRAPEZOID)
        for edge in edge_order:
            p1 = edge
            p2 = (edge + 1) % n
            self.add_point(p1)
            self.add_point(p2)
            self.add_segment(p1, p2)
            if trace is not None:
                trace.append((edge, list(self.node_stack), len(self.nodes)))
        N = self.nodes
        mountains = []  # [p1, p2, list]
        def m_add(key, p1, p2):
            found = None
            for m in mountains:
                if m[0] == N[key].point1 and m[1] == N[key].point2:
                    found = m
            if found is None:
                found = [_req(N[key].point1), _req(N[key].point2), []]
                mountains.append(found)
            found[2].append(p1)
            found[2].append(p2)
        for item in range(len(N)):
            nd = N[item]
            if nd.type != TRAPEZOID:
                continue
            if nd.child1 is not None:
                c1 = N[nd.child1]
                if not (c1.crumb == c1.child2):
                    continue
            else:
                continue
            point1, point2 = _req(nd.point1), _req(nd.point2)
            child1, child2 = _req(nd.child1), _req(nd.child2)
            if point1 == N[child2].point1 and point2 == N[child2].point2:
                m_add(child1, point1, point2)
            elif point1 == N[child1].point1 and point2 == N[child1].point2:
                m_add(child2, point1, point2)
            else:
                m_add(child1, point1, point2)
                m_add(child2, point1, point2)
        for m in mountains:
            L = m[2]
            # insertion sort, stable
            for i in range(1, len(L)):
                j = i
                while j > 0 and self.point_is_above(L[j], L[j - 1]):
                    L[j], L[j - 1] = L[j - 1], L[j]
                    j -= 1
            while len(L) > 2:
                p1, p2, p3 = len(L) - 2, len(L) - 1, 0
                progressed = False
                for item in range(1, len(L)):
                    if L[p1] == L[p2]:
                        del L[p1]
                        progressed = True
                        break
                    if L[p2] == L[p3]:
                        del L[p2]
                        progressed = True
                        break
                    if self.push_triangle_if_acute(L[p2], L[p1], L[p3], emit):
                        del L[p2]
                        progressed = True
                        break
                    p1, p2, p3 = p2, p3, item
                if not progressed:
                    self.stuck = True
                    break
        return mountains


def triangulate_ids(points, edge_order, atan2=None):
    """Returns (list of emitted point ids, status string)."""
    t = Triangulation(atan2)
    out = []
    try:
        t.create_polygon(points, edge_order, out.append)
    except NullUnwrap:
        return out, "null_unwrap"
    return out, "stuck" if t.stuck else "ok"
