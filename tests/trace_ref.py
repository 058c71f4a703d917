"""A THIRD restatement of the reference's estimator -- test infrastructure, written from the Rust source and sharing no
code with oracle/oracle.cpp or the CUDA kernels -- for the analytic subset: bare unit Sphere and Plane objects, Point /
Directional / Ambient lights, constant environment, any Material.  Pure Python on IEEE doubles (`math` = the C libm the
oracle links), Philox4x32-10 on Python integers.  tests/test_oracle.py demands BIT-EQUALITY of its images with
oracle.cpp's: two independent readings of the same source can only agree to the last bit if both evaluate the same
expressions in the same order, which is what "faithful restatement" means.

    Renderer::get_color / trace_ray / sample_lights / get_closest_hit   src/renderer.rs:131-220
    Camera::cast_ray                                                     src/camera.rs:64-81
    Sphere::intersect, Plane::intersect                                   src/shape/sphere.rs:13-45, src/shape/plane.rs:17-32
    Light::illuminate                                                     src/light.rs:23-32
    Material::bsdf, sample_f, local_to_world                              src/material.rs:125-324
    rand 0.8 / rand_distr 0.4 draws on the repo's Philox stream (seed, pixel, sample)
"""
import math

INF = float("inf")
EPSILON = 1e-12      # renderer.rs:14
FIREFLY_CLAMP = 100.0  # renderer.rs:15
M32 = 0xFFFFFFFF


# ---------------------------------------------------------------- vectors (nalgebra-glm, left to right) --------
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def neg(a): return (-a[0], -a[1], -a[2])
def mul(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def div(a, s): return (a[0] / s, a[1] / s, a[2] / s)
def cmul(a, b): return (a[0] * b[0], a[1] * b[1], a[2] * b[2])
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def normalize(a): return div(a, math.sqrt(dot(a, a)))
def sign_positive(x): return math.copysign(1.0, x) > 0.0   # f64::is_sign_positive: -0.0 is negative
def signum(x): return x if x != x else math.copysign(1.0, x)
def is_normal(x): return x == x and abs(x) != INF and abs(x) >= 2.2250738585072014e-308


def powi(a, b):  # Rust f64::powi = compiler-rt __powidf2
    r = 1.0
    while True:
        if b & 1:
            r *= a
        b //= 2
        if b == 0:
            return r
        a *= a


# ---------------------------------------------------------------- Philox4x32-10 + rand 0.8 semantics ------------
def philox_block(block, pixel, s_lo, s_hi, k0, k1):
    c0, c1, c2, c3 = block, pixel, s_lo, s_hi
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class Rng:
    """64-bit draw 4b + w = (word w of Philox block b) << 32 | (word w of block b | 2^31); counter = (block, pixel, sample)."""

    def __init__(self, seed, pixel, sample):
        self.k0, self.k1 = seed & M32, (seed >> 32) & M32
        self.pixel, self.s_lo, self.s_hi = pixel, sample & M32, (sample >> 32) & M32
        self.block = 0
        self.pending = []

    def next_u64(self):
        if not self.pending:
            hi = philox_block(self.block, self.pixel, self.s_lo, self.s_hi, self.k0, self.k1)
            lo = philox_block(self.block | 0x80000000, self.pixel, self.s_lo, self.s_hi, self.k0, self.k1)
            self.block += 1
            self.pending = [(h << 32) | l for h, l in zip(hi, lo)]
        return self.pending.pop(0)

    def gen(self):  # Standard f64: 53 bits
        return float(self.next_u64() >> 11) * (1.0 / 9007199254740992.0)

    def u52(self):  # the [1, 2) mantissa trick of UniformFloat, minus 1
        return float(self.next_u64() >> 12) * (1.0 / 4503599627370496.0)

    def gen_range(self, lo, hi):  # UniformFloat::sample_single
        scale = hi - lo
        while True:
            res = (1.0 + self.u52()) * scale + (lo - scale)
            if res < hi:
                return res

    def pm1(self):  # Uniform::new(-1, 1)
        return self.u52() * 2.0 + (-1.0)

    def gen_bool(self, p):  # Bernoulli: one draw always
        v = self.next_u64()
        if p >= 1.0:
            return True
        return v < int(p * 18446744073709551616.0)

    def unit_disc(self):
        while True:
            x, y = self.pm1(), self.pm1()
            if x * x + y * y <= 1.0:
                return x, y

    def unit_circle(self):
        while True:
            x1, x2 = self.pm1(), self.pm1()
            s = x1 * x1 + x2 * x2
            if s < 1.0:
                return (x1 * x1 - x2 * x2) / s, 2.0 * x1 * x2 / s


# ---------------------------------------------------------------- scene description -----------------------------
class Mat:
    def __init__(self, color, index, roughness, metallic, emittance, transparent):
        self.color, self.index, self.roughness = tuple(color), index, roughness
        self.metallic, self.emittance, self.transparent = metallic, emittance, transparent


class Sphere:  # the unit sphere at the origin
    def intersect(self, o, d, t_min, rec):
        a = dot(d, d)
        b = dot(d, o)
        c = dot(o, o) - 1.0
        disc = b * b - a * c
        if not sign_positive(disc):
            return False
        sd = math.sqrt(disc)
        t = (-b - sd) / a
        if t < t_min:
            t = (-b + sd) / a
            if t < t_min:
                return False
        if t < rec[0]:
            rec[0] = t
            rec[1] = normalize(add(o, mul(d, t)))
            return True
        return False


class Plane:
    def __init__(self, normal, value):
        self.normal, self.value = tuple(normal), value

    def intersect(self, o, d, t_min, rec):
        cosine = dot(self.normal, d)
        if abs(cosine) < 1e-8:
            return False
        time = (self.value - dot(self.normal, o)) / cosine
        if time >= t_min and time < rec[0]:
            rec[0] = time
            rec[1] = mul(neg(normalize(self.normal)), signum(cosine))
            return True
        return False


# ---------------------------------------------------------------- Material ---------------------------------------
def lerp3(a, b, t): return add(mul(a, 1.0 - t), mul(b, t))


def bsdf(m, n, wo, wi):
    ndwi, ndwo = dot(n, wi), dot(n, wo)
    wi_out, wo_out = sign_positive(ndwi), sign_positive(ndwo)
    if not m.transparent and (not wi_out or not wo_out):
        return (0.0, 0.0, 0.0)
    one = (1.0, 1.0, 1.0)
    m2 = m.roughness * m.roughness
    if wi_out == wo_out:
        h = normalize(add(wi, wo))
        wodh, ndh = dot(wo, h), dot(n, h)
        nh2 = powi(ndh, 2)
        d = math.exp((nh2 - 1.0) / (m2 * nh2)) / (m2 * math.pi * nh2 * nh2)
        if not wi_out and math.sqrt(1.0 - wodh * wodh) * m.index > 1.0:
            f = one
        else:
            f0 = powi((m.index - 1.0) / (m.index + 1.0), 2)
            f0 = lerp3((f0, f0, f0), m.color, m.metallic)
            f = add(f0, mul(sub(one, f0), powi(1.0 - wodh, 5)))
        g = min(ndwi * ndh, ndwo * ndh)
        g = (2.0 * g) / wodh
        g = min(g, 1.0)
        spec = div(mul(mul(f, d), g), 4.0 * ndwo * ndwi)
        if m.transparent:
            return spec
        return add(spec, div(cmul(sub(one, f), m.color), math.pi))
    eta = m.index if wo_out else 1.0 / m.index
    h = normalize(add(mul(wi, eta), wo))
    widh, wodh, ndh = dot(wi, h), dot(wo, h), dot(n, h)
    nh2 = powi(ndh, 2)
    d = math.exp((nh2 - 1.0) / (m2 * nh2)) / (m2 * math.pi * nh2 * nh2)
    f0 = powi((m.index - 1.0) / (m.index + 1.0), 2)
    f0 = lerp3((f0, f0, f0), m.color, m.metallic)
    f = add(f0, mul(sub(one, f0), powi(1.0 - abs(widh), 5)))
    g = min(abs(ndwi * ndh), abs(ndwo * ndh))
    g = (2.0 * g) / abs(wodh)
    g = min(g, 1.0)
    btdf = mul(div(mul(mul(sub(one, f), d), g), powi(eta * widh + wodh, 2)), abs(widh * wodh / (ndwi * ndwo)))
    return cmul(btdf, m.color)


def local_to_world(n, v):
    ns = normalize((n[1], -n[0], 0.0)) if is_normal(n[0]) else normalize((0.0, -n[2], n[1]))
    nss = cross(n, ns)
    # mat3 rows (ns.x nss.x n.x / ...) times v = ns*v.x + nss*v.y + n*v.z, summed left to right per component
    return (ns[0] * v[0] + nss[0] * v[1] + n[0] * v[2], ns[1] * v[0] + nss[1] * v[1] + n[1] * v[2], ns[2] * v[0] + nss[2] * v[1] + n[2] * v[2])


def sample_f(m, n, wo, rng):
    m2 = m.roughness * m.roughness
    f0 = powi((m.index - 1.0) / (m.index + 1.0), 2)
    f = (1.0 - m.metallic) * f0 + m.metallic * ((m.color[0] + m.color[1] + m.color[2]) / 3.0)
    f = f * (1.0 - 0.2) + 1.0 * 0.2
    eta = m.index if dot(wo, n) > 0.0 else 1.0 / m.index

    def beckmann():
        theta = math.atan(math.sqrt(m2 * -math.log(rng.gen())))
        sin_t, cos_t = math.sin(theta), math.cos(theta)
        x, y = rng.unit_circle()
        return local_to_world(n, (x * sin_t, y * sin_t, cos_t))

    def beckmann_pdf(h):
        cos_t = abs(dot(h, n))
        sin_t = math.sqrt(1.0 - cos_t * cos_t)
        return (1.0 / (math.pi * m2 * powi(cos_t, 3))) * math.exp(-powi(sin_t / cos_t, 2) / m2)

    if rng.gen_bool(f):
        h = beckmann()
        wi = neg(sub(wo, mul(h, 2.0 * dot(h, wo))))  # -reflect_vec(wo, h), reflect = i - 2 (n.i) n
    elif not m.transparent:
        x, y = rng.unit_disc()
        z = math.sqrt(1.0 - x * x - y * y)
        wi = local_to_world(n, (x, y, z))
    else:
        h = beckmann()
        cos_to = dot(h, wo)
        wo_perp = sub(wo, mul(h, cos_to))
        wi_perp = div(neg(wo_perp), eta)
        sin2 = dot(wi_perp, wi_perp)
        if sin2 > 1.0:
            return None
        cos_ti = math.sqrt(1.0 - sin2)
        wi = add(mul(h, -signum(cos_to) * cos_ti), wi_perp)
    p = 0.0
    h = normalize(add(wi, wo))
    p += f * beckmann_pdf(h) / (4.0 * abs(dot(h, wo)))
    if not m.transparent:
        p += (1.0 - f) * max(dot(wi, n), 0.0) * (1.0 / math.pi)
    elif sign_positive(dot(wo, n)) != sign_positive(dot(wi, n)):
        h = normalize(add(mul(wi, eta), wo))
        hwo, hwi = dot(h, wo), dot(h, wi)
        p += (1.0 - f) * beckmann_pdf(h) * (abs(hwo) / powi(eta * hwi + hwo, 2))
    return wi, p


# ---------------------------------------------------------------- Renderer ---------------------------------------
class Renderer:
    def __init__(self, objects, lights, env, eye, direction, up, fov, width, height, max_bounces, exposure_value=0.0):
        """objects: [(shape, Mat)]; lights: [("point", color, location) | ("directional", color, direction) | ("ambient", color)]"""
        self.objects, self.lights, self.env = objects, lights, tuple(env)
        self.eye, self.direction, self.up, self.fov = tuple(eye), tuple(direction), tuple(up), fov
        self.width, self.height, self.max_bounces, self.ev = width, height, max_bounces, exposure_value
        self.segments = 0

    def closest_hit(self, o, d):
        rec = [INF, (0.0, 0.0, 0.0)]
        hit = None
        for shape, mat in self.objects:
            if shape.intersect(o, d, EPSILON, rec):
                hit = mat
        return (rec, hit) if hit is not None else None

    def sample_lights(self, m, pos, n, wo, rng):
        color = (0.0, 0.0, 0.0)
        for l in self.lights:
            if l[0] == "ambient":
                color = add(color, cmul(l[1], m.color))
                continue
            if l[0] == "point":
                disp = sub(l[2], pos)
                ln = math.sqrt(dot(disp, disp))
                intensity, wi, dist = div(l[1], ln * ln), div(disp, ln), ln
            else:
                intensity, wi, dist = l[1], neg(normalize(l[2])), INF
            h = self.closest_hit(pos, wi)
            if h is None or h[0][0] > dist:
                color = add(color, mul(cmul(bsdf(m, n, wo, wi), intensity), dot(wi, n)))
        return color

    def trace_ray(self, o, d, depth, rng):
        self.segments += 1
        h = self.closest_hit(o, d)
        if h is None:
            return self.env
        (t, n), m = h
        pos = add(o, mul(d, t))
        wo = neg(normalize(d))
        color = mul(m.color, m.emittance)
        color = add(color, self.sample_lights(m, pos, n, wo, rng))
        if depth < self.max_bounces:
            s = sample_f(m, n, wo, rng)
            if s is not None:
                wi, pdf = s
                f = bsdf(m, n, wo, wi)
                ind = mul(mul(cmul(f, self.trace_ray(pos, wi, depth + 1, rng)), 1.0 / pdf), abs(dot(wi, n)))
                color = (color[0] + min(ind[0], FIREFLY_CLAMP), color[1] + min(ind[1], FIREFLY_CLAMP), color[2] + min(ind[2], FIREFLY_CLAMP))
        return color

    def cast_ray(self, x, y):
        d = 1.0 / math.tan(self.fov / 2.0)
        right = normalize(cross(self.direction, self.up))
        new_dir = add(add(mul(self.direction, d), mul(right, x)), mul(self.up, y))
        return self.eye, normalize(new_dir)

    def get_color(self, x, y, iterations, seed, first_sample=0):
        dim = float(max(self.width, self.height))
        xn = (float(2 * x + 1) - float(self.width)) / dim
        yn = (float(2 * (self.height - y) - 1) - float(self.height)) / dim
        color = (0.0, 0.0, 0.0)
        for s in range(iterations):
            rng = Rng(seed, y * self.width + x, first_sample + s)
            dx = rng.gen_range(-1.0 / dim, 1.0 / dim)
            dy = rng.gen_range(-1.0 / dim, 1.0 / dim)
            o, d = self.cast_ray(xn + dx, yn + dy)
            color = add(color, self.trace_ray(o, d, 0, rng))
        return mul(div(color, float(iterations)), 2.0 ** self.ev)

    def sample(self, iterations, seed, first_sample=0):
        return [self.get_color(x, y, iterations, seed, first_sample) for y in range(self.height) for x in range(self.width)]


def from_api(scene, camera, width, height, max_bounces, exposure_value=0.0) -> "Renderer":
    """Build the restatement's scene from the host mirror's objects (rpt_b200.api): bare Sphere / Plane shapes only."""
    from rpt_b200 import _capi as capi

    objs = []
    for o in scene.objects:
        name = type(o.shape).__name__
        m = o.mat
        mat = Mat([float(c) for c in m.color], m.index, m.roughness, m.metallic, m.emittance, m.transparent)
        if name == "Sphere":
            objs.append((Sphere(), mat))
        elif name == "Plane":
            objs.append((Plane([float(c) for c in o.shape.normal], float(o.shape.value)), mat))
        else:
            raise ValueError("trace_ref handles bare spheres and planes only, not " + name)
    lights = []
    for l in scene.lights:
        c = tuple(float(v) for v in l.color)
        v = tuple(float(x) for x in l.vec)
        if l.kind == capi.LIGHT_POINT:
            lights.append(("point", c, v))
        elif l.kind == capi.LIGHT_DIRECTIONAL:
            lights.append(("directional", c, v))
        elif l.kind == capi.LIGHT_AMBIENT:
            lights.append(("ambient", c))
        else:
            raise ValueError("trace_ref has no object lights")
    assert scene.environment.hdri is None
    return Renderer(objs, lights, [float(c) for c in scene.environment.color], [float(c) for c in camera.eye],
                    [float(c) for c in camera.direction], [float(c) for c in camera.up], float(camera.fov), width, height, max_bounces,
                    exposure_value)
