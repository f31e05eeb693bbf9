"""``smartpy.objfunctions`` surface (smartpy/objfunctions.py:20-24).

In the batch path this constraint is evaluated inside the kernel (scores column 'GW');
this scalar form is kept for callers of the reference API."""


def groundwater_constraint(evaluation, simulation):
    lower, upper = evaluation[0] - 0.1, evaluation[0] + 0.1
    return 1.0 if (lower <= simulation[0]) and (simulation[0] <= upper) else 0.0
