"""TEST INFRASTRUCTURE -- fixed-column PDB ``ATOM`` parser exposing the biopandas column names the
reference relies on (``src/inference_rigid.py:77-82, 148-156``).  Off the hot path: used only to
regenerate graph fixtures from the reference's shipped test PDBs."""
import pandas as pd


class PandasPdb:
    def __init__(self):
        self.df = {}

    def read_pdb(self, path):
        rows = []
        with open(path, 'r') as fh:
            for line_idx, line in enumerate(fh):
                if not line.startswith('ATOM'):
                    continue
                line = line.rstrip('\n').ljust(80)
                rows.append({
                    'record_name': 'ATOM',
                    'atom_number': int(line[6:11]),
                    'atom_name': line[12:16].strip(),
                    'alt_loc': line[16].strip(),
                    'residue_name': line[17:20].strip(),
                    'chain_id': line[21].strip(),
                    'residue_number': int(line[22:26]),
                    'insertion': line[26].strip(),
                    'x_coord': float(line[30:38]),
                    'y_coord': float(line[38:46]),
                    'z_coord': float(line[46:54]),
                    'occupancy': float(line[54:60]) if line[54:60].strip() else 1.0,
                    'b_factor': float(line[60:66]) if line[60:66].strip() else 0.0,
                    'element_symbol': line[76:78].strip(),
                    'line_idx': line_idx,
                })
        self.df['ATOM'] = pd.DataFrame(rows)
        return self
