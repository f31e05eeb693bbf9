"""numpy restatement of the scoring half of the hot path (TEST INFRASTRUCTURE).

The reference scores each sample at ``smartpy/montecarlo/montecarlo.py:193-209`` by calling
four functions of the third-party package ``spotpy`` (``spotpy.objectivefunctions``,
constraint ``>=1.5.14`` at the reference's ``setup.py:56-58``, no pin; its source is NOT
under /root/reference and the package is not installable offline), plus
``smartpy/objfunctions.py:20-24``.  The published definitions are restated here:

    nashsutcliffe  1 - sum((e - s)^2) / sum((e - mean(e))^2)
    kge            1 - sqrt((r-1)^2 + (alpha-1)^2 + (beta-1)^2),
                   r = corrcoef(e, s)[0, 1], alpha = std(s)/std(e) (ddof 0), beta = sum(s)/sum(e)
    pbias          100 * sum(s - e) / sum(e)
    rmse           sqrt(mean((e - s)^2))

Parity pin: the only artefact of the reference that fixes these semantics is
``examples/out/ExampleDaily/ExampleDaily.SMART.lhs`` (10 samples, float32); see
tests/test_oracle_golden.py::test_scores_match_reference_lhs_file.  Beyond float32
precision the scoring parity is UNPINNED (stated in DESIGN.md).
"""
import numpy as np

SCORE_NAMES = ['NSE', 'KGE', 'KGEc', 'KGEa', 'KGEb', 'PBias', 'RMSE', 'GW']


def nashsutcliffe(evaluation, simulation):
    e, s = np.asarray(evaluation, dtype=np.float64), np.asarray(simulation, dtype=np.float64)
    return 1 - np.sum((e - s) ** 2) / np.sum((e - np.mean(e)) ** 2)


def kge(evaluation, simulation, return_all=False):
    e, s = np.asarray(evaluation, dtype=np.float64), np.asarray(simulation, dtype=np.float64)
    cc = np.corrcoef(e, s)[0, 1]
    alpha = np.std(s) / np.std(e)
    beta = np.sum(s) / np.sum(e)
    k = 1 - np.sqrt((cc - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)
    return (k, cc, alpha, beta) if return_all else k


def pbias(evaluation, simulation):
    e, s = np.asarray(evaluation, dtype=np.float64), np.asarray(simulation, dtype=np.float64)
    return 100 * (float(np.sum(s - e)) / float(np.sum(e)))


def rmse(evaluation, simulation):
    e, s = np.asarray(evaluation, dtype=np.float64), np.asarray(simulation, dtype=np.float64)
    return np.sqrt(np.mean((e - s) ** 2))


def groundwater_constraint(evaluation, simulation):
    """objfunctions.py:20-24."""
    if (evaluation[0] - 0.1 <= simulation[0]) and (simulation[0] <= evaluation[0] + 0.1):
        return 1.0
    return 0.0


def objectivefunction(simulation, evaluation, gw_constraint=None):
    """MonteCarlo.objectivefunction (montecarlo.py:193-209).

    simulation = (discharge[n_report], [gw]); evaluation = (obs[n_report], [gw_constraint]).
    Returns [NSE, KGE, KGEc, KGEa, KGEb, PBias, RMSE] (+ [GW] when the constraint is set,
    montecarlo.py:71-74 -- note the reference tests truthiness, so a constraint of 0.0 is off).
    """
    obs = np.asarray(evaluation[0], dtype=np.float64)
    mask = ~np.isnan(obs)
    e = obs[mask]
    s = np.asarray(simulation[0], dtype=np.float64)[mask]
    o1 = nashsutcliffe(e, s)
    o2, o2c, o2a, o2b = kge(e, s, return_all=True)
    o3 = pbias(e, s)
    o4 = rmse(e, s)
    res = [o1, o2, o2c, o2a, o2b, o3, o4]
    if gw_constraint:
        res.append(groundwater_constraint([gw_constraint], simulation[1]))
    return res


def score_members(discharge, gw, obs, gw_constraint=None):
    """Rows of discharge[N, n_report] -> scores[N, 8] (GW column NaN when no constraint)."""
    discharge = np.atleast_2d(discharge)
    out = np.full((discharge.shape[0], 8), np.nan)
    for i in range(discharge.shape[0]):
        r = objectivefunction((discharge[i], [gw[i]]), (obs, [gw_constraint]), gw_constraint)
        out[i, :len(r)] = r
    return out
