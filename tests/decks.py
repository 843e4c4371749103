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


# ---- examples/RT3D.py (the deck of tests/cases/testRT.py and of BASELINE config 4) ---------------
def rt_mesh(npts, two_d=True):
    """examples/RT3D.py:28-43."""
    ff = 2 * float(np.pi) * float(npts - 1) / npts
    if two_d:
        dom = ((2.8 * float(np.pi), int(npts * 1.4), False), (ff, npts, True), (ff, 1, True))
    else:
        dom = ((3.0 * float(np.pi), int(npts * 1.5), False), (ff, npts, True), (ff, npts, True))
    return "\n".join("%sdom = (0.0, %r, %d, periodic=%s)" % (c, L, n, p) for c, (L, n, p) in zip("xyz", dom))


def rt_xbar(pysim, data):
    """examples/RT3D.py:73-79 `meanXto3d`, as written there: the y-z mean as a 3-D field, through
    `pysim.PyMPI.yzsum`, `pysim.emptyScalar()` and `pysim.mesh.indices`."""
    meanX = pysim.PyMPI.yzsum(data) / (pysim.PyMPI.ny * pysim.PyMPI.nz)
    tmp = pysim.emptyScalar()
    for i in range(tmp.shape[0]):
        ii = int(pysim.mesh.indices[0].data[i, 0, 0])
        tmp[i, :, :] = meanX[ii]
    return tmp


def RT_PARMS(npts):
    """examples/RT3D.py:46-65 (the dictionary is applied to the deck text, pyranda.py:219-228)."""
    return {"gx": -0.01, "CPh": 1.4, "CPl": 1.4, "CVh": 1.0, "CVl": 1.0, "mwH": 3.0, "mwL": 1.0, "Runiv": 1.0,
            "waveLength": 4, "rho_l": 1.0, "rho_h": 3.0, "delta": 2.0 * np.pi / npts * 4}


# examples/RT3D.py:86-185
RT_EOM = """
ddt(:rhoYh:)  =  -ddx(:rhoYh:*:u: - :Jx:)    - ddy(:rhoYh:*:v: - :Jy:)   - ddz(:rhoYh:*:w: - :Jz:)
ddt(:rhoYl:)  =  -ddx(:rhoYl:*:u: + :Jx:)    - ddy(:rhoYl:*:v: + :Jy:)   - ddz(:rhoYl:*:w: + :Jz:)
ddt(:rhou:)   =  -ddx(:rhou:*:u: - :tauxx:)  - ddy(:rhou:*:v: - :tauxy:) - ddz(:rhou:*:w: - :tauxz:) + :rho:*gx
ddt(:rhov:)   =  -ddx(:rhov:*:u: - :tauxy:)  - ddy(:rhov:*:v: - :tauyy:) - ddz(:rhov:*:w: - :tauyz:)
ddt(:rhow:)   =  -ddx(:rhow:*:u: - :tauxz:)  - ddy(:rhow:*:v: - :tauyz:) - ddz(:rhow:*:w: - :tauzz:)
ddt(:Et:)     =  -ddx( (:Et: - :tauxx:)*:u: - :tauxy:*:v: - :tauxz:*:w: - :tx:*:kappa:) - ddy( (:Et: - :tauyy:)*:v: -:tauxy:*:u: - :tauyz:*:w: - :ty:*:kappa:) - ddz( (:Et: - :tauzz:)*:w: - :tauxz:*:u: - :tauyz:*:v: - :tz:*:kappa:) + :rho:*gx*:u:
:rhoYh:     =  fbar( :rhoYh:  )
:rhoYl:     =  fbar( :rhoYl:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:rhow:      =  fbar( :rhow: )
:Et:        =  fbar( :Et:   )
:mybar: = xbar(:rho:*:u:*:u:)
:rho:       = :rhoYh: + :rhoYl:
:Yh:        =  :rhoYh: / :rho:
:Yl:        =  :rhoYl: / :rho:
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:w:         =  :rhow: / :rho:
:cv:        = :Yh:*CVh + :Yl:*CVl
:cp:        = :Yh:*CPh + :Yl:*CPl
:gamma:     = :cp:/:cv:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v:) ) * ( :gamma: - 1.0 )
:mw:        = 1.0 / ( :Yh: / mwH + :Yl: / mwL )
:R:         = Runiv / :mw:
:T:         = :p: / (:rho: * :R: )
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
:Yx:        =  ddx(:Yh:)
:Yy:        =  ddy(:Yh:)
:Yz:        =  ddz(:Yh:)
:enst:      = sqrt( (:uy:-:vx:)**2 + (:uz: - :wx:)**2 + (:vz:-:wy:)**2 )
:S:         = sqrt( :ux:*:ux: + :vy:*:vy: + :wz:*:wz: + .5*((:uy:+:vx:)**2 + (:uz: + :wx:)**2 + (:vz:+:wy:)**2) )
:mu:        = 1.0e-4 * gbar( abs(ring(:S:  )) ) * :rho:
:beta:      = 7.0e-2 * gbar( abs(ring(:div:)) * :rho: )
:Dsgs:      =  1.0e-4 * ring(:Yh:)
:Ysgs:      =  1.0e2  * (abs(:Yh:) - 1.0 + abs(1.0-:Yh: ) )*gridLen**2
:adiff:     =  gbar( :rho:*numpy.maximum(:Dsgs:,:Ysgs:) / :dt: )
:Jx:        =  :adiff:*:Yx:
:Jy:        =  :adiff:*:Yy:
:Jz:        =  :adiff:*:Yz:
:taudia:    =  (:beta:-2./3.*:mu:) *:div: - :p:
:tauxx:     =  2.0*:mu:*:ux:   + :taudia:
:tauyy:     =  2.0*:mu:*:vy:   + :taudia:
:tauzz:     =  2.0*:mu:*:wz:   + :taudia:
:tauxy:     = :mu:*(:uy:+:vx:)
:tauxz:     = :mu:*(:uz:+:wx:)
:tauyz:     = :mu:*(:vz:+:wz:)
[:tx:,:ty:,:tz:] = grad(:T:)
:kappa:     = 1.0e-3 * gbar( ring(:T:)* :rho:*:cv:/(:T: * :dt: ) )
:cs:        = sqrt( :p: / :rho: * :gamma: )
:dt:        = dt.courant(:u:,:v:,:w:,:cs:)*1.0
:dt:        = numpy.minimum(:dt:,0.2 * dt.diff(:beta:,:rho:))
:dt:        = numpy.minimum(:dt:,0.2 * dt.diff(:mu:,:rho:))
:dt:        = numpy.minimum(:dt:,0.2 * dt.diff(:adiff:,:rho:))
bc.const(['Yh'],['xn'],1.0)
bc.const(['Yh'],['x1'],0.0)
bc.const(['Yl'],['x1'],1.0)
bc.const(['Yl'],['xn'],0.0)
bc.extrap(['rho','Et'],['x1','xn'])
bc.const(['u','v','w'],['x1','xn'],0.0)
:leftBC:  = 1.0 - 0.5 * (1.0 + tanh( (meshx-1.0) / .1 ) )
:rightBC: = 0.5 * (1.0 + tanh( (meshx-8.0) / .1 ) )
:BC: = numpy.maximum(:leftBC:,:rightBC:)
:u: = gbar(:u:)*:BC: + :u:*(1.0 - :BC:)
:rho: = gbar(:rho:)*:BC: + :rho:*(1.0 - :BC:)
:Et: = gbar(:Et:)*:BC: + :Et:*(1.0 - :BC:)
:rhoYh: = :rho:*:Yh:
:rhoYl: = :rho:*:Yl:
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:rhow: = :rho:*:w:
:mix:  = 4.0*:Yh:*:Yl:
"""

# examples/RT3D.py:192-222
RT_IC = """
:gamma:= 5./3.
p0     = 1.0
At     = ( (rho_h - rho_l)/(rho_h + rho_l) )
u0     = sqrt( abs(gx*At/waveLength) ) * .1
:Yl: = .5 * (1.0-tanh( sqrt(pi)*( meshx - 1.4*pi ) / delta ))
:Yh: = 1.0 - :Yl:
:p:  += p0
wgt = 4*:Yh:*:Yl:
:v: *= 0.0
:w: *= 0.0
:u: = wgt * gbar( (random3D()-0.5)*u0 )
:rho:       = rho_h * :Yh: + rho_l * :Yl:
:cv:        = :Yh:*CVh + :Yl:*CVl
:cp:        = :Yh:*CPh + :Yl:*CPl
:gamma:     = :cp:/:cv:
:mw:        = 1.0 / ( :Yh: / mwH + :Yl: / mwL )
:R:         = Runiv / :mw:
:T:         = :p: / (:rho: * :R: )
:rhoYh: = :rho:*:Yh:
:rhoYl: = :rho:*:Yl:
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:rhow: = :rho:*:w:
:Et:  = :p: / (:gamma:-1.0) + 0.5*:rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
"""


# ---- examples/KH.py (tests/cases/test2deuler.py: KH-2d-64) ----------------------------------------
def kh_mesh(npts):
    """examples/KH.py:48-52."""
    L = 1.0 * (npts - 1.0) / npts
    return "xdom = (0.0, %r, %d, periodic=True)\nydom = (0.0, %r, %d, periodic=True)" % (L, npts, L, npts)


KH_GAMMA = 5. / 3.
KH_EOM_PARMS = {"gamma": KH_GAMMA, "R0": 1.0, "cv": 1.0 / (1.0 - 1.0 / KH_GAMMA) - 1.0}  # examples/KH.py:8-21,104
KH_IC_PARMS = {"gamma": KH_GAMMA, "u1": 0.5, "u2": -0.5, "p0": 2.5, "rho1": 1.0, "rho2": 2.0, "L": 0.025, "Vmean": 0.0}

# examples/KH.py:60-101
KH_EOM = """
ddt(:rho:)  =  -ddx(:rho:*:u:)            - ddy(:rho:*:v:)
ddt(:rhou:) =  -ddx(:rhou:*:u: - :tauxx:) - ddy(:rhou:*:v: - :tauxy:)
ddt(:rhov:) =  -ddx(:rhov:*:u: - :tauxy:) - ddy(:rhov:*:v: - :tauyy:)
ddt(:Et:)   =  -ddx( (:Et: - :tauxx:)*:u: - :tauxy:*:v:  - :tx:*:kappa: )  - ddy( (:Et: - :tauyy:)*:v: -:tauxy:*:u:- :ty:*:kappa: )
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:Et:        =  fbar( :Et:   )
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v:) ) * ( gamma - 1.0 )
:ux:        =  ddx(:u:)
:vy:        =  ddy(:v:)
:div:       =  :ux: + :vy:
:uy:        =  ddy(:u:)
:vx:        =  ddx(:v:)
:enst:      = sqrt( (:uy:-:vx:)**2 )
:tke:       = :rho:*(:u:*:u: + :v:*:v: )
:S:         = sqrt( :ux:*:ux: + :vy:*:vy:  + .5*((:uy:+:vx:)**2  ) )
:mu:        =  gbar( abs(ring(:S:  )) ) * :rho: * 1.0e-4
:beta:      =  gbar( abs(ring(:div:)) * :rho: )  * 7.0e-3
:taudia:    =  (:beta:-2./3.*:mu:) *:div: - :p:
:tauxx:     =  2.0*:mu:*:ux:   + :taudia:
:tauyy:     =  2.0*:mu:*:vy:   + :taudia:
:tauxy:     = :mu:*(:uy:+:vx:)
:T:         = :p: / (:rho: * R0 )
[:tx:,:ty:,:tz:] = grad(:T:)
:kappa:     = gbar( ring(:T:)* :rho:*cv/(:T: * :dt: ) ) * 1.0e-3
:cs:  = sqrt( :p: / :rho: * gamma )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)*1.0
:dt: = numpy.minimum(:dt:,0.2 * dt.diff(:beta:,:rho:))
:dt: = numpy.minimum(:dt:,0.2 * dt.diff(:mu:,:rho:))
"""

# examples/KH.py:108-127
KH_IC = """
Um = (u1-u2)/2.0
rhoM = (rho1-rho2)/2.0
:u: =                     u1-Um*exp( -(meshy-.75)/L)
:u: = where( meshy < .75, u2+Um*exp( -(.75-meshy)/L) , :u: )
:u: = where( meshy < .50, u2+Um*exp( (-meshy+.25)/L) , :u: )
:u: = where( meshy < .25, u1-Um*exp(  (meshy-.25)/L) , :u: )
:v: = 0.01*sin( 4.*pi*meshx ) + Vmean
:rho: =                     rho1-rhoM*exp( -(meshy-.75)/L)
:rho: = where( meshy < .75, rho2+rhoM*exp( -(.75-meshy)/L) , :rho: )
:rho: = where( meshy < .50, rho2+rhoM*exp( (-meshy+.25)/L) , :rho: )
:rho: = where( meshy < .25, rho1-rhoM*exp(  (meshy-.25)/L) , :rho: )
:p: += p0
:Et: = :p:/( gamma - 1.0 ) + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * gamma )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)*.1
"""


# ---- examples/euler.py, dim = 2, problem = 'sod' (tests/cases/test2deuler.py: euler-2d-64) --------
def euler2d_mesh(npts):
    """examples/euler.py:36-47."""
    Lp = float(np.pi) * 2.0 * (npts - 1.0) / npts
    return {"x1": [0.0, 0.0, 0.0], "xn": [Lp, Lp, Lp], "nn": [npts, npts, 1], "periodic": [False, False, True]}


# examples/euler.py:53-74
EULER2D_EOM = """
ddt(:rho:)  =  -ddx(:rho:*:u:)                  - ddy(:rho:*:v:)
ddt(:rhou:) =  -ddx(:rhou:*:u: + :p: - :tau:)   - ddy(:rhou:*:v:)
ddt(:rhov:) =  -ddx(:rhov:*:u:)                 - ddy(:rhov:*:v: + :p: - :tau:)
ddt(:Et:)   =  -ddx( (:Et: + :p: - :tau:)*:u: ) - ddy( (:Et: + :p: - :tau:)*:v: )
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:Et:        =  fbar( :Et:   )
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v:) ) * ( :gamma: - 1.0 )
:div:       =  ddx(:u:) + ddy(:v:)
:beta:      =  gbar(abs(ring(:div:))) * :rho: * 7.0e-2
:tau:       =  :beta:*:div:
bc.extrap(['rho','Et'],['x1','xn','y1','yn'])
bc.const(['u','v'],['x1','xn','y1','yn'],0.0)
"""

# examples/euler.py:88-113
EULER2D_IC = """
rad = sqrt( (meshx-pi)**2  +  (meshy-pi)**2 )
:gamma: = 1.4
:Et:  = gbar( where( rad < pi/2.0, 1.0/(:gamma:-1.0) , .1 /(:gamma:-1.0) ) )
:rho: = gbar( where( rad < pi/2.0, 1.0    , .125 ) )
"""


# ---- examples/cylinder.py (tests/cases/testCylinder.py: cylinder-2d-32 / -64) ---------------------
def cylinder_mesh(npts):
    """examples/cylinder.py:33-41."""
    Lp = float(np.pi) * 2.0 * (npts - 1.0) / npts
    return {"x1": [0.0, 0.0, 0.0], "xn": [Lp, Lp, Lp], "nn": [npts, npts, 1], "periodic": [False, False, False]}


# examples/cylinder.py:51-57,104,130: Mach-2 inflow; the deck text is patched with str() of these
CYL_MACH = 2.0
CYL_U0 = float(np.sqrt(1.0 / 1.0 * 1.4)) * CYL_MACH

# examples/cylinder.py:61-103
CYLINDER_EOM = """
ddt(:rho:)  =  -ddx(:rho:*:u:)                  - ddy(:rho:*:v:)
ddt(:rhou:) =  -ddx(:rhou:*:u: + :p: - :tau:)   - ddy(:rhou:*:v:)
ddt(:rhov:) =  -ddx(:rhov:*:u:)                 - ddy(:rhov:*:v: + :p: - :tau:)
ddt(:Et:)   =  -ddx( (:Et: + :p: - :tau:)*:u: - :tx:*:kappa:) - ddy( (:Et: + :p: - :tau:)*:v: - :ty:*:kappa: )
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:Et:        =  fbar( :Et:   )
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v:) ) * ( :gamma: - 1.0 )
:T:         = :p: / (:rho: * :R: )
:div:       =  ddx(:u:) + ddy(:v:)
:beta:      =  gbar( ring(:div:) * :rho: ) * 7.0e-2
:tau:       = :beta:*:div:
[:tx:,:ty:,:tz:] = grad(:T:)
:kappa:     = gbar( ring(:T:)* :rho:*:cv:/(:T: * :dt: ) ) * 1.0e-3
[:u:,:v:,:w:] = ibmV( [:u:,:v:,:w:], :phi:, [:gx:,:gy:,:gz:], [:u1:,:u2:,0.0] )
:rho: = ibmS( :rho: , :phi:, [:gx:,:gy:,:gz:] )
:p:   = ibmS( :p:   , :phi:, [:gx:,:gy:,:gz:] )
bc.extrap(['rho','p','u'],['xn'])
bc.const(['u'],['x1','y1','yn'],u0)
bc.const(['v'],['x1','xn','y1','yn'],0.0)
bc.const(['rho'],['x1','y1','yn'],rho0)
bc.const(['p'],['x1','y1','yn'],p0)
:Et:  = :p: / ( :gamma: - 1.0 )  + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
:dtB: = 0.2* dt.diff(:beta:,:rho:)
:dt: = numpy.minimum(:dt:,:dtB:)
:umag: = sqrt( :u:*:u: + :v:*:v: )
""".replace('u0', str(CYL_U0)).replace('p0', str(1.0)).replace('rho0', str(1.0))

# examples/cylinder.py:110-129
CYLINDER_IC = """
:gamma: = 1.4
:R: = 1.0
:cp: = :R: / (1.0 - 1.0/:gamma: )
:cv: = :cp: - :R:
rad = sqrt( (meshx-pi)**2  +  (meshy-pi)**2 )
:phi: = rad - pi/4.0
:rho: = 1.0 + 3d()
:p:  =  1.0 + 3d() #exp( -(meshx-1.5)**2/.25**2)*.1
:u: = where( :phi:>0.5, mach * sqrt( :p: / :rho: * :gamma:) , 0.0 )
:u: = gbar( gbar( :u: ) )
:v: = 0.0 + 3d()
:Et: = :p:/( :gamma: - 1.0 ) + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)*.1
[:gx:,:gy:,:gz:] = grad( :phi: )
:gx: = gbar( :gx: )
:gy: = gbar( :gy: )
""".replace('mach', str(CYL_MACH))


# ---- examples/cylinder_curv.py (tests/cases/testCylinder.py: cylinder_curved-2d-64; BASELINE config 5) ----
def zoom_mesh_1d(npts, x1, xn, xa, xb, tt, dxf):
    """examples/meshTest.py:7-55 zoomMesh_solve: spacing dxf inside [xa, xb], blended by tanh into a
    coarse spacing found by a secant iteration so that the last node lands on xn."""
    def zoom(dxc):
        x = np.zeros((npts))
        x[0] = x1
        dx = dxc
        for i in range(1, npts):
            x[i] = x[i - 1] + dx
            xmin = (xa + xb) / 2.0
            xh = (xb - xa) / 2.0
            xpr = np.abs(x[i] - xmin)
            w = 0.5 * (np.tanh((xpr - xh) / tt) + 1.0)
            dx = w * dxc + (1.0 - w) * dxf
        return x
    nmin = (xb - xa) / dxf
    dxc = ((xn - x1) - (xb - xa)) / (npts - nmin)
    delta, f1, cnt = 1.0001, xn, 1
    while (abs(f1) > .01 * xn) and cnt < 100:
        f1 = zoom(dxc)[-1] - xn
        f2 = zoom(dxc * delta)[-1] - xn
        dxc = dxc * (1.0 - f1 / (f2 - f1) * (delta - 1.0))
        cnt += 1
    return zoom(dxc)


def cylinder_curv_mesh(npts):
    """examples/cylinder_curv.py:33-62."""
    Lp = float(np.pi) * 2.0 * (npts - 1.0) / npts
    dxf = 4 * Lp / float(npts) * .3
    xS = zoom_mesh_1d(npts, -2. * Lp, 2. * Lp, -2., 2., 1.0, dxf)
    return {"coordsys": 3, "function": lambda i, j, k: (xS[i], xS[j], 0.0), "periodic": [False, False, True],
            "periodicGrid": False, "x1": [-2 * Lp, -2 * Lp, 0.0], "xn": [2 * Lp, 2 * Lp, Lp], "nn": [npts, npts, 1]}


# examples/cylinder_curv.py:83-119
CYLINDER_CURV_EOM = """
ddt(:rho:)  =  -div(:rho:*:u:,  :rho:*:v:)
ddt(:rhou:) =  -div(:rhou:*:u: + :p: - :tau:, :rhou:*:v:)
ddt(:rhov:) =  -div(:rhov:*:u:, :rhov:*:v: + :p: - :tau:)
ddt(:Et:)   =  -div( (:Et: + :p: - :tau:)*:u: - :tx:*:kappa:, (:Et: + :p: - :tau:)*:v: - :ty:*:kappa: )
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:Et:        =  fbar( :Et:   )
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v:) ) * ( :gamma: - 1.0 )
:T:         = :p: / (:rho: * :R: )
:div:       =  div(:u:,:v:)
:beta:      =  gbar( ring(:div:) * :rho: ) * 7.0e-3
:tau:       =  :beta: * :div:
[:tx:,:ty:,:tz:] = grad(:T:)
:kappa:     = gbar( ring(:T:)* :rho:*:cv:/(:T: * :dt: ) ) * 1.0e-3
[:u:,:v:,:w:] = ibmV( [:u:,:v:,:w:], :phi:, [:gx:,:gy:,:gz:], [:u1:,:u2:,0.0] )
:rho: = ibmS( :rho: , :phi:, [:gx:,:gy:,:gz:] )
:p:   = ibmS( :p:   , :phi:, [:gx:,:gy:,:gz:] )
bc.extrap(['rho','p','u'],['xn'])
bc.const(['u'],['x1','y1','yn'],u0)
bc.const(['v'],['x1','xn','y1','yn'],0.0)
bc.const(['rho'],['x1','y1','yn'],rho0)
bc.const(['p'],['x1','y1','yn'],p0)
:Et:  = :p: / ( :gamma: - 1.0 )  + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
:dtB: = 0.2* dt.diff(:beta:,:rho:)
:dt: = numpy.minimum(:dt:,:dtB:)
:umag: = sqrt( :u:*:u: + :v:*:v: )
""".replace('u0', str(CYL_U0)).replace('p0', str(1.0)).replace('rho0', str(1.0))

# examples/cylinder_curv.py:126-146
CYLINDER_CURV_IC = """
:gamma: = 1.4
:R: = 1.0
:cp: = :R: / (1.0 - 1.0/:gamma: )
:cv: = :cp: - :R:
rad = sqrt( meshx**2  +  meshy**2 )
:phi: = rad - pi/4.0
:rho: = 1.0 + 3d()
:p:  =  1.0 + 3d() #exp( -(meshx-1.5)**2/.25**2)*.1
:u: = where( :phi:>0.5, mach * sqrt( :p: / :rho: * :gamma:) , 0.0 )
:u: = gbar( gbar( :u: ) )
:v: = 0.0 + 3d()
:Et: = :p:/( :gamma: - 1.0 ) + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
[:gx:,:gy:,:gz:] = grad( :phi: )
:gx: = gbar( :gx: )
:gy: = gbar( :gy: )
""".replace('mach', str(CYL_MACH))


# ---- examples/cylinder_curv2.py (tests/cases/testCylinder.py: cylinder_omesh-2d-64) ----------------
def cylinder_omesh(npts):
    """examples/cylinder_curv2.py:33-60: an O-grid around the cylinder, periodic in the angle."""
    Lp = float(np.pi) * 2.0 * (npts - 1.0) / npts
    Ri, Rf, NX, NY = 1.0, 10.0, npts, npts

    def cyl(i, j, k):
        theta = float(i) / float(NX) * 2.0 * np.pi
        r = float(j) / float(NY - 1) * (Rf - Ri) + Ri
        return r * np.cos(theta), r * np.sin(theta), 0.0
    return {"coordsys": 3, "function": cyl, "periodic": [True, False, False], "periodicGrid": False,
            "x1": [-2 * Lp, -2 * Lp, 0.0], "xn": [2 * Lp, 2 * Lp, Lp], "nn": [NX, NY, 1]}


OMESH_MACH = 1.5
OMESH_U0 = float(np.sqrt(1.0 / 1.0 * 1.4)) * OMESH_MACH  # examples/cylinder_curv2.py:69-75

# examples/cylinder_curv2.py:78-107
OMESH_EOM = """
ddt(:rho:)  =  -div(:rho:*:u:,  :rho:*:v:)
ddt(:rhou:) =  -div(:rhou:*:u: + :p: - :tau:, :rhou:*:v:)
ddt(:rhov:) =  -div(:rhov:*:u:, :rhov:*:v: + :p: - :tau:)
ddt(:Et:)   =  -div( (:Et: + :p: - :tau:)*:u: , (:Et: + :p: - :tau:)*:v:  )
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:Et:        =  fbar( :Et:   )
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v:) ) * ( :gamma: - 1.0 )
:div:       =  div(:u:,:v:)
:beta:      =  gbar( ring(:div:) * :rho: ) * 7.0e-2
:tau:       = :beta:*:div:
bc.extrap(['u','v','rho','p'],['yn'])
bc.const(['u'],['yn'],u0)
bc.const(['v'],['yn'],0.0)
bc.extrap(['rho','p'],['y1'])
bc.slip([ ['u','v']  ],['y1'])
:Et:  = :p: / ( :gamma: - 1.0 )  + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
:dtB: = dt.diff(:beta:,:rho:)
:umag: = sqrt( :u:*:u: + :v:*:v: )
""".replace('u0', str(OMESH_U0)).replace('p0', str(1.0)).replace('rho0', str(1.0))

# examples/cylinder_curv2.py:113-128
OMESH_IC = """
:gamma: = 1.4
:R: = 1.0
:cp: = :R: / (1.0 - 1.0/:gamma: )
:cv: = :cp: - :R:
rad = sqrt( meshx**2  +  meshy**2 )
:rho: = 1.0 + 3d()
:p:  =  1.0 + 3d() #exp( -(meshx-1.5)**2/.25**2)*.1
:u: = mach * sqrt( :p: / :rho: * :gamma:)
:v: = 0.0 + 3d()
:Et: = :p:/( :gamma: - 1.0 ) + .5*:rho:*(:u:*:u: + :v:*:v:)
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
""".replace('mach', str(OMESH_MACH))


# ---- a box with symmetry planes (no reference example uses meshOptions['symmetric']; this deck exists to
# run the SYMM operators -- even closures everywhere, the odd first derivative inside div -- through
# the interpreter on both backends) ------------------------------------------------------------------
def symm_box_mesh():
    return {"x1": [0.0, 0.0, 0.0], "xn": [1.0, 1.0, 1.0], "nn": [32, 24, 20], "periodic": [False, False, False],
            "symmetric": [[True, True], [True, False], [False, True]]}


SYMM_EOM = """
ddt(:phi:) = -div(:phi:*:u:, :phi:*:v:, :phi:*:w:) + 0.01*lap(:phi:)
:phi: = fbar(:phi:)
[:gx:,:gy:,:gz:] = grad(:phi:)
:r: = ring(:phi:) + gbar(:phi:) + 1.0e-4*(dd4x(:phi:) + dd4y(:phi:) + dd4z(:phi:))
:dt: = dt.courant(:u:,:v:,:w:,:c:)
"""

SYMM_IC = """
:phi: = 1.0 + 0.2*cos(pi*meshx)*cos(2.0*pi*meshy)*cos(pi*meshz)
:u: = 0.3*sin(pi*meshx)*cos(pi*meshy)
:v: = -0.3*cos(pi*meshx)*sin(pi*meshy)
:w: = 0.1*sin(2.0*pi*meshz)
:c: = 1.0 + 3d()
:dt: = dt.courant(:u:,:v:,:w:,:c:)
"""


# examples/3Dadvect.py:14-20,33-97 (BASELINE configs[0]: 3-D advection of a density bump, periodic, the
# artificial viscosities switched off by their 0.0e-4 factors), strings as written there
def advect3d_mesh(npts):
    return tgv_mesh(npts)


ADVECT3D_EOM = """
# Primary Equations of motion here
ddt(:rho:)  =  -ddx(:rho:*:u:)                  - ddy(:rho:*:v:)                  - ddz(:rho:*:w:)
ddt(:rhou:) =  -ddx(:rhou:*:u: + :p: - :tauxx:) - ddy(:rhou:*:v: - :tauxy:)       - ddz(:rhou:*:w: - :tauxz:)
ddt(:rhov:) =  -ddx(:rhov:*:u: - :tauxy:)       - ddy(:rhov:*:v: + :p: - :tauyy:) - ddz(:rhov:*:w: - :tauyz:)
ddt(:rhow:) =  -ddx(:rhow:*:u: - :tauxz:)       - ddy(:rhow:*:v: - :tauyz:)       - ddz(:rhow:*:w: + :p: - :tauzz:)
ddt(:Et:)   =  -ddx( (:Et: - :tauxx:)*:u: - :tauxy:*:v: - :tauxz:*:w: ) - ddy( (:Et: - :tauyy:)*:v: -:tauxy:*:u: - :tauyz:*:w:) - ddz( (:Et: - :tauzz:)*:w: - :tauxz:*:u: - :tauyz:*:v: )
# Conservative filter of the EoM
:rho:       =  fbar( :rho:  )
:rhou:      =  fbar( :rhou: )
:rhov:      =  fbar( :rhov: )
:rhow:      =  fbar( :rhow: )
:Et:        =  fbar( :Et:   )
# Update the primatives and enforce the EOS
:u:         =  :rhou: / :rho:
:v:         =  :rhov: / :rho:
:w:         =  :rhow: / :rho:
:p:         =  ( :Et: - .5*:rho:*(:u:*:u: + :v:*:v: + :w:*:w:) ) * ( :gamma: - 1.0 )
# Artificial bulk viscosity 
:ux:        =  ddx(:u:)
:vy:        =  ddy(:v:)
:wz:        =  ddz(:w:)
:div:       =  :ux: + :vy: + :wz:
# Remaining cross derivatives
:uy:        =  ddy(:u:)
:uz:        =  ddz(:u:)
:vx:        =  ddx(:v:)
:vz:        =  ddz(:v:)
:wy:        =  ddy(:w:)
:wx:        =  ddx(:w:)
:enst:      = sqrt( (:uy:-:vx:)**2 + (:uz: - :wx:)**2 + (:vz:-:wy:)**2 )
:tke:       = :rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:S:         = sqrt( :ux:*:ux: + :vy:*:vy: + :wz:*:wz: + .5*((:uy:+:vx:)**2 + (:uz: + :wx:)**2 + (:vz:+:wy:)**2) )
:mu:        =  gbar( abs(ring(:S:  )) ) * :rho: * 0.0e-4
:beta:      =  gbar( abs(ring(:div:)) * :rho: )  * 0.0e-4
:taudia:    =  (:beta:-2./3.*:mu:) *:div:
:tauxx:     =  2.0*:mu:*:ux:   + :taudia: - :p:
:tauyy:     =  2.0*:mu:*:vy:   + :taudia: - :p:
:tauzz:     =  2.0*:mu:*:wz:   + :taudia: - :p:
:tauxy:     = :mu:*(:uy:+:vx:) + :taudia:
:tauxz:     = :mu:*(:uz:+:wx:) + :taudia:
:tauyz:     = :mu:*(:vz:+:wz:) + :taudia:
:cs:  = sqrt( :p: / :rho: * :gamma: )
:dt: = dt.courant(:u:,:v:,:w:,:cs:)*.5
:dt: = numpy.minimum(:dt:,0.2 * dt.diff(:beta:,:rho:))
:dt: = numpy.minimum(:dt:,0.2 * dt.diff(:mu:,:rho:))
"""

ADVECT3D_IC = """
:gamma: = 1.4
u0 = 1.0
p0 = 1.0
rho0 = 1.0
L = 1.0
rad = sqrt( (meshx-pi)**2 + (meshy-pi)**2 + (meshz-pi)**2 )
:w: +=  u0
:p:  += p0 
:rho: += rho0*(1.0+ .01* (1.0+tanh( -(rad-.5)/(.5) )) )
:rhou: = :rho:*:u:
:rhov: = :rho:*:v:
:rhow: = :rho:*:w:
:Et:  = :p: / (:gamma:-1.0) + 0.5*:rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:cs:  = sqrt( :p: / :rho: * :gamma: )
:tke: = :rho:*(:u:*:u: + :v:*:v: + :w:*:w:)
:dt: = dt.courant(:u:,:v:,:w:,:cs:)
"""
