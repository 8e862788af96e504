class FileIO(object):
    @staticmethod
    def load_data_set(file, rec_type="graph"):
        with open(file) as f:
            next(f)
            return [[int(a), int(b), 1.0] for a, b in (line.split("\t")[:2] for line in f if line.strip())]

    @staticmethod
    def write_file(path, name, lines):  # something the B200 package does not replace
        with open(path + name, "w") as f:
            f.writelines(lines)
