"""The ten-parameter SMART set -- same public surface as the reference's
``smartpy/parameters.py:22-104`` (``names``, ``ranges``, ``values`` and the two setters,
same error messages).  The order of ``names`` is the column order of every ``params[N, 10]``
array handed to the CUDA kernel (include/smart_b200.h)."""
import csv

_TABLE = (
    # name, lower, upper            (typical ranges, parameters.py:27-38)
    ('T', 0.9, 1.1),        # rainfall aerial correction coefficient
    ('C', 0.0, 1.0),        # evaporation decay parameter
    ('H', 0.0, 0.3),        # quick runoff coefficient
    ('D', 0.0, 1.0),        # drain flow parameter
    ('S', 0.0, 0.013),      # soil outflow coefficient
    ('Z', 15.0, 150.0),     # effective soil depth [mm]
    ('SK', 1.0, 240.0),     # surface routing parameter [h]
    ('FK', 48.0, 1440.0),   # interflow routing parameter [h]
    ('GK', 1200.0, 4800.0),  # groundwater routing parameter [h]
    ('RK', 1.0, 96.0),      # river routing parameter [h]
)


class Parameters(object):
    def __init__(self):
        #: Return the SMART model parameter names as a `list`.
        self.names = [row[0] for row in _TABLE]
        #: Return the typical SMART model parameter ranges as a `dict`.
        self.ranges = {row[0]: (row[1], row[2]) for row in _TABLE}
        #: Return the set of SMART model parameter values as a `dict`.
        self.values = dict()

    def set_parameters_with_file(self, file_location):
        """Assign the SMART model parameters values using a CSV file with the columns
        ``PAR_NAME,PAR_VALUE`` (e.g. ``examples/in/ExampleDaily/ExampleDaily.parameters``)."""
        found = dict()
        try:
            with open(file_location, 'r', encoding='utf8') as handle:
                for record in csv.DictReader(handle):
                    if record['PAR_NAME'] in self.names:
                        found[record['PAR_NAME']] = float(record['PAR_VALUE'])
        except KeyError:
            raise Exception("There is 'PAR_NAME' or 'PAR_VALUE' column in {}.".format(file_location))
        except ValueError:
            raise Exception("There is at least one incorrect parameter value in {}.".format(file_location))
        except IOError:
            raise Exception("There is no parameters file at {}.".format(file_location))

        missing = [name for name in self.names if name not in found]
        if missing:
            raise Exception("The parameter {} is not available in the "
                            "parameters file at {}.".format(missing[0], file_location))
        for name in self.names:
            self.values[name] = found[name]

    def set_parameters_with_dict(self, dictionary):
        """Assign the SMART model parameters values using a dictionary keyed by ``names``."""
        for name in self.names:
            if name not in dictionary:
                raise Exception("The parameter {} is not available in the dictionary provided.")
            self.values[name] = dictionary[name]

    def as_row(self, values=None):
        """Values in ``names`` order (the kernel's parameter layout)."""
        values = self.values if values is None else values
        return [values[name] for name in self.names]
