"""Known answers the reference's own tests hold for the OPTIMISER-DRIVEN calls on this path.  They are
trajectory dependent (max_iter = 4 of Optim.LBFGS + HagerZhang), so reproducing them pins the
restatement of the third-party optimiser (oracle/optim_lbfgs.py) and the device driver
(csrc/optimizer.cu).  Literals copied with their file:line (paths under the reference tree)."""
import numpy as np

# test/localization/max_localize.jl:45-56  max_localize(p, model) with DEFAULT tolerances, atol 1e-7
MAXLOC_DEFAULT = dict(Omega=6.374823673444644, OmegaI=5.812709709242578, OmegaOD=0.5621139641912166,
                      OmegaTilde=0.5621139642020658)

# test/localization/disentangle.jl:56-97  disentangle(p, model; max_iter=4), atol 1e-7
DISENTANGLE_IT4 = dict(
    Omega=18.13021209358854, OmegaI=11.672692248601113, OmegaOD=6.341211889610597,
    OmegaD=0.11630795537683003, OmegaTilde=6.457519844987427,
    omega=np.array([1.6720402167866713, 2.4604144313827963, 2.466947090855194, 2.4657348106032355,
                    1.672054382451226, 2.463760873012476, 2.471821111520987, 2.457439176975955]),
    r=np.array([[1.349540244604139, 1.3494591284925146, 1.3494989831767787],
                [1.3483994786002111, 1.3489677557627042, 1.3492175157250734],
                [1.349148846101609, 1.349041631018211, 1.3487479885732723],
                [1.348490312952771, 1.3492694250911146, 1.348720943361702],
                [-0.00013820836974004684, -5.965595567273871e-5, -9.848597753393869e-5],
                [0.0009659756244854475, 0.00035987176361155105, 0.00018872798640355847],
                [0.00026663589947916155, 0.0004041165538806068, 0.0006513524118075078],
                [0.0009449317726510505, 0.00011446347253599799, 0.000661121718585481]]))

# test/localization/constrain_center/max_localize.jl:46-96  CenterSpreadPenalty(r0 = 0, lambda = 10),
# max_localize(p, model; max_iter=4) from the parallel-transport gauge, atol 1e-7
CC_MAXLOC_IT4 = dict(
    r0=np.zeros((4, 3)), lam=10.0,
    Omega=10.79175007994904, OmegaI=5.8127097092426245, OmegaOD=4.912331468930677,
    OmegaD=0.06670890177573785, OmegaTilde=4.979040370706415,
    omega=np.array([1.974272148108204, 2.9590704528882212, 2.957966420231007, 2.9004410587216065]),
    r=np.array([[-0.0006371414769181565, 0.0006811595685799503, 0.0300072108960674],
                [-0.011666353819322479, -0.0026638301666946987, -0.04279486032530962],
                [0.0057784441053171055, -0.0036622860182758767, -0.03176818666470274],
                [0.00550194992572012, 0.005763601542074829, 0.04839785532365684]]),
    Omegac=0.06337765901009804, Omegat=10.855127738959137,
    omegac=np.array([0.009013026333805437, 0.01974599872857372, 0.0105602043912133, 0.024058429556505577]))

# test/localization/constrain_center/disentangle.jl:43-120  r0 = 4 x (1.3494,)*3 + 4 x 0, lambda = 10,
# disentangle(p, model; max_iter=4), atol 1e-7
CC_DISENTANGLE_IT4 = dict(
    r0=np.array([[1.34940] * 3] * 4 + [[0.0] * 3] * 4), lam=10.0,
    Omega=18.13023065207267, OmegaI=11.672695160371777, OmegaOD=6.341223286965608,
    OmegaD=0.11631220473528359, OmegaTilde=6.457535491700892,
    omega=np.array([1.6720494155597025, 2.460415452388486, 2.4669444682373154, 2.4657363554531777,
                    1.6720635891342623, 2.4637620130565874, 2.4718181805922725, 2.457441177650863]),
    r=np.array([[1.3493218953976438, 1.3493613581763424, 1.3493454033039636],
                [1.3487523304656146, 1.3491260454857306, 1.3493415966363236],
                [1.349320637608541, 1.3491584356558521, 1.348968696047302],
                [1.348788920937121, 1.349359045780774, 1.3489543861927666],
                [8.250336965638709e-5, 3.423661589664105e-5, 5.730256938007102e-5],
                [0.0006279664190044885, 0.00023781998769869842, 5.6624178504504735e-5],
                [8.189201302412916e-5, 0.0002630773899438576, 0.0004277217819759191],
                [0.0006247774390737991, 3.728918078729418e-5, 0.0004318068660840961]]),
    Omegac=2.6352789712419767e-5, Omegat=18.130257004862383)


def check_spread(sp, ref, atol=1e-7):
    """sp: dict with Omega, OmegaI, OmegaOD, OmegaD, OmegaTilde, omega, r (oracle / api naming)"""
    get = (lambda k: sp[k]) if isinstance(sp, dict) else (lambda k: getattr(sp, k))
    for key in ("Omega", "OmegaI", "OmegaOD", "OmegaD", "OmegaTilde"):
        if key in ref:
            assert abs(get(key) - ref[key]) < atol, (key, get(key), ref[key])
    if "omega" in ref:
        assert np.abs(np.asarray(get("omega")) - ref["omega"]).max() < atol
    if "r" in ref:
        assert np.abs(np.asarray(get("r")) - ref["r"]).max() < atol
