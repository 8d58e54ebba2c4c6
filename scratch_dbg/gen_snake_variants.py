import sys; sys.path.insert(0,'oracle'); sys.path.insert(0,'oracle/shims')
import numpy as np
import elastica as ea
from elastica.external_forces import MuscleTorques
class Sim(ea.BaseSystemCollection, ea.Constraints, ea.Forcing, ea.Damping, ea.Contact, ea.CallBacks): pass
def run(mu_scale, tag, out):
    sim = Sim(); n=50; L=0.35; r=L*0.011; E=1e6; dt=8e-6; period=2.0
    rod = ea.CosseratRod.straight_rod(n, np.zeros(3), np.array([0,0,1.]), np.array([0,1.,0]), L, r, 1000, youngs_modulus=E, shear_modulus=E/1.5)
    sim.append(rod)
    sim.dampen(rod).using(ea.AnalyticalLinearDamper, damping_constant=1e-4, time_step=dt)
    sim.add_forcing_to(rod).using(ea.GravityForces, acc_gravity=np.array([0.0,-9.80665,0.0]))
    b = np.array([5.4791206e-03, -1.2224312e-03, 7.1719582e-03, 3.9473604e-03, -8.1164530e-03, 9.5124468e-03], dtype=np.float32)
    kw = np.float32(2*np.pi)/np.float32(2.4028492)
    sim.add_forcing_to(rod).using(MuscleTorques, base_length=L, b_coeff=b, period=period, wave_number=kw, phase_shift=0.0,
        rest_lengths=rod.rest_lengths, ramp_up_time=period, direction=np.array([0,1.,0]), with_spline=True)
    mu = L/(period*period*9.80665*0.1)*mu_scale
    plane = ea.Plane(plane_origin=np.array([0,-r,0.]), plane_normal=np.array([0,1.,0]))
    sim.append(plane)
    sim.detect_contact_between(rod, plane).using(ea.RodPlaneContactWithAnisotropicFriction, k=1.0, nu=1e-6, slip_velocity_tol=1e-8,
        static_mu_array=np.zeros(3), kinetic_mu_array=np.array([mu,1.5*mu,2*mu]))
    sim.finalize()
    st = ea.PositionVerlet(); t = np.float64(0.0); done=0
    for tgt in (1, 10, 100, 2083, 6000):
        for _ in range(tgt-done): t = st.step(sim, t, dt)
        done = tgt
        for k in ("position_collection","velocity_collection","director_collection","omega_collection"):
            out[f"{tag}/s{tgt}/{k}"] = getattr(rod,k).copy()
out = {}
run(0.0, "nofric", out); run(1.0, "fric", out)
np.savez("scratch_dbg/snake_variants.npz", **out); print("ok")
