import math


def _norm(deg):
    # AngleDeg normalisation: fmod into (-360, 360) if outside, then fold into [-180, 180].
    if deg < -360.0 or 360.0 < deg:
        deg = math.fmod(deg, 360.0)
    if deg < -180.0:
        deg += 360.0
    if deg > 180.0:
        deg -= 360.0
    return deg


class AngleDeg:
    def __init__(self, deg=0.0):
        self._d = _norm(float(deg.degree() if isinstance(deg, AngleDeg) else deg))

    def degree(self):
        return self._d

    def abs(self):
        return math.fabs(self._d)

    def __sub__(self, other):
        o = other.degree() if isinstance(other, AngleDeg) else float(other)
        return AngleDeg(self._d - o)

    def __add__(self, other):
        o = other.degree() if isinstance(other, AngleDeg) else float(other)
        return AngleDeg(self._d + o)

    def __repr__(self):
        return f"{self._d}"

    @staticmethod
    def atan2_deg(y, x):
        # (*) librcsc tests for the exact zero vector; pyrusgeom may use an epsilon.  Immaterial
        # away from the origin; the oracle documents the same choice.
        if x == 0.0 and y == 0.0:
            return 0.0
        return math.degrees(math.atan2(y, x))


class Vector2D:
    def __init__(self, x=0.0, y=0.0):
        self._x = float(x)
        self._y = float(y)

    def x(self):
        return self._x

    def y(self):
        return self._y

    def abs_x(self):
        return math.fabs(self._x)

    def abs_y(self):
        return math.fabs(self._y)

    def r(self):
        return math.sqrt(self._x * self._x + self._y * self._y)

    def th(self):
        return AngleDeg(AngleDeg.atan2_deg(self._y, self._x))

    def dist(self, other):
        dx, dy = self._x - other._x, self._y - other._y
        return math.sqrt(dx * dx + dy * dy)

    def __sub__(self, o):
        return Vector2D(self._x - o._x, self._y - o._y)

    def __add__(self, o):
        return Vector2D(self._x + o._x, self._y + o._y)

    def __repr__(self):
        return f"({self._x}, {self._y})"

    @staticmethod
    def from_polar(r, ang):
        d = ang.degree() if isinstance(ang, AngleDeg) else float(ang)
        rad = math.radians(d)
        return Vector2D(r * math.cos(rad), r * math.sin(rad))
