class PrettyTable(object):
    def __init__(self, *args, **kwargs):
        self.field_names, self.rows = [], []

    def add_row(self, row):
        self.rows.append(row)

    def __str__(self):
        return "\n".join(str(r) for r in [self.field_names] + self.rows)
