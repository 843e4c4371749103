"""Problem decks used by the tests and benchmarks: the equation / initial-condition strings of the
reference's examples, verbatim in content (examples/TaylorGreen.py:26-108, examples/advection.py)."""
import numpy as np


def tgv_mesh(npts):
    L = str(2 * float(np.pi) * float(npts - 1) / npts)
    return "\n".join("%sdom = (0.0, %s, %d, periodic=True)" % (c, L, npts) for c in "xyz")


# examples/TaylorGreen.py:40-87 (including its `:tauyz:` line as written there)
TGV_EOM = """
ddt(:rho:)  =  -ddx(:rho:*:u:)            - ddy(:rho:*:v:)               - ddz(:rho:*:w:)
ddt(:rhou:) =  -ddx(:rhou:*:u: - :tauxx:) - ddy(:rhou:*:v: - :tauxy:)    - ddz(:rhou:*:w: - :tauxz:)
ddt(:rhov:) =  -ddx(:rhov:*:u: - :tauxy:) - ddy(:rhov:*:v: - :tauyy:)    - ddz(:rhov:*:w: - :tauyz:)
ddt(:rhow:) =  -ddx(:rhow:*:u: - :tauxz:) - ddy(:rhow:*:v: - :tauyz:)    - ddz(:rhow:*:w: - :tauzz:)
ddt(:Et:)   =  -ddx( (:Et: - :tauxx:)*:u: - :tauxy:*:v: - :tauxz:*:w: )  - ddy( (:Et: - :tauyy:)*:v: -:tauxy:*:u: - :tauyz:*:w:) - ddz( (:Et: - :tauzz:)*:w: - :tauxz:*:u: - :tauyz:*:v: )
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:rhow:      =  fbar( :rhow: )
:Et:        =  fbar( :Et:   )
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:w:         =  :rhow: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v: + :w:*:w:) ) * ( :gamma: - 1.0 )
:ux:        =  ddx(:u:)
:vy:        =  ddy(:v:)
:wz:        =  ddz(:w:)
:div:       =  :ux: + :vy: + :wz:
:uy:        =  ddy(:u:)
:uz:        =  ddz(:u:)
:vx:        =  ddx(:v:)
:vz:        =  ddz(:v:)
:wy:        =  ddy(:w:)
:wx:        =  ddx(:w:)
:enst:      = sqrt( (:uy:-:vx:)**2 + (:uz: - :wx:)**2 + (:vz:-:wy:)**2 )
:tke:       = :rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:S:         = sqrt( :ux:*:ux: + :vy:*:vy: + :wz:*:wz: + .5*((:uy:+:vx:)**2 + (:uz: + :wx:)**2 + (:vz:+:wy:)**2) )
:mu:        =  gbar( abs(ring(:S:  )) ) * :rho: * 1.0e-4
:beta:      =  gbar( abs(ring(:div:)) * :rho: )  * 7.0e-3
:taudia:    =  (:beta:-2./3.*:mu:) *:div: - :p:
:tauxx:     =  2.0*:mu:*:ux:   + :taudia:
:tauyy:     =  2.0*:mu:*:vy:   + :taudia:
:tauzz:     =  2.0*:mu:*:wz:   + :taudia:
:tauxy:     = :mu:*(:uy:+:vx:)
:tauxz:     = :mu:*(:uz:+:wx:)
:tauyz:     = :mu:*(:vz:+:wz:)
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)*1.0
:dt: = numpy.minimum(:dt:,0.2 * dt.diff(:beta:,:rho:))
:dt: = numpy.minimum(:dt:,0.2 * dt.diff(:mu:,:rho:))
"""

# examples/TaylorGreen.py:92-108
TGV_IC = """
:gamma: = 1.4
u0 = 1.0
p0 = 100.0
rho0 = 1.0
L = 1.0
:u: =  u0*sin(meshx/L)*cos(meshy/L)*cos(meshz/L)
:v: = -u0*cos(meshx/L)*sin(meshy/L)*cos(meshz/L)
:w: = 0.0*:u:
:p:  = p0 + rho0/16.0*( ( cos(2.*meshx/L) + cos(2.*meshy/L) ) * ( cos(2.*meshz/L) + 2.0 ) - 2.0 )
:rho: = rho0 + 0.0*:u:
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:rhow: = :rho:*:w:
:Et:  = :p: / (:gamma:-1.0) + 0.5*:rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:cs:  = sqrt( :p: / :rho: * :gamma: )
:tke: = :rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
"""


def run_tgv(ss, tstop=None, nsteps=None, cfl=0.5):
    """examples/TaylorGreen.py:110-160: returns the enstrophy ratio history."""
    ss.EOM(TGV_EOM)
    ss.setIC(TGV_IC)
    time = 0.0
    dt = ss.variables["dt"] * cfl
    enst0 = ss.B.sum3D(ss.variables["enst"])
    enst = 1.0
    n = 0
    while (tstop is not None and time < tstop) or (nsteps is not None and n < nsteps):
        time = ss.rk4(time, dt)
        dt = ss.variables["dt"] * cfl
        enst = ss.B.sum3D(ss.variables["enst"]) / enst0
        n += 1
    return enst, time, n


# A bounded (non-periodic) 2-D advection deck with the BC package, in the style of the reference's
# boundary-driven examples (examples/advection.py + pyrandaBC usage in examples/cylinder.py:70-90).
def bc_mesh(n):
    return "\n".join(["xdom = (0.0, 1.0, %d)" % n, "ydom = (0.0, 1.0, %d)" % n, "zdom = (0.0, 1.0, 1)"])


BC_EOM = """
ddt(:phi:)  =  - :c: * ddx(:phi:) - 0.5 * :c: * ddy(:phi:)
:phi:       =  fbar(:phi:)
bc.extrap(['phi'],['xn','yn'])
bc.const(['phi'],['x1'],0.0)
bc.extrap(['phi'],['y1'],order=1)
:grad2:     =  ddx(:phi:)*ddx(:phi:) + ddy(:phi:)*ddy(:phi:)
bc.field('grad2',['x1'],:phi:)
"""

BC_IC = """
:c:   = 1.0
:phi: = exp(-((meshx-0.4)**2 + (meshy-0.5)**2)/0.02)
"""
