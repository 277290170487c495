"""BOP result files: the writer scripts/infer.py uses for its pose estimates, with the row format of
bop_toolkit_lib.inout.save_bop_results(version='bop19')
(/root/reference/external/bop_toolkit/bop_toolkit_lib/inout.py:265-294), and the matching reader (:220-262) that the
BOP evaluation scripts use -- kept here so that the tests can read back what infer.py wrote."""
import numpy as np

HEADER = 'scene_id,im_id,obj_id,score,R,t,time'


def save_bop_results(path, results, version='bop19'):
    """results: list of {'scene_id','im_id','obj_id','score','R' [3,3],'t' [3,1], optional 'time'} (infer.py:492-503)."""
    if version != 'bop19':
        raise ValueError('Unknown version of BOP results.')
    lines = [HEADER]
    for res in results:
        lines.append('{scene_id},{im_id},{obj_id},{score},{R},{t},{time}'.format(
            scene_id=res['scene_id'], im_id=res['im_id'], obj_id=res['obj_id'], score=res['score'],
            R=' '.join(map(str, np.asarray(res['R']).flatten().tolist())),
            t=' '.join(map(str, np.asarray(res['t']).flatten().tolist())),
            time=res['time'] if 'time' in res else -1))
    with open(path, 'w') as f:
        f.write('\n'.join(lines))


def load_bop_results(path, version='bop19'):
    if version != 'bop19':
        raise ValueError('Unknown version of BOP results.')
    results = []
    with open(path, 'r') as f:
        for line_id, line in enumerate(f, 1):
            if line_id == 1 and HEADER in line:
                continue
            elems = line.split(',')
            if len(elems) != 7:
                raise ValueError('A line does not have 7 comma-sep. elements: {}'.format(line))
            results.append({
                'scene_id': int(elems[0]), 'im_id': int(elems[1]), 'obj_id': int(elems[2]), 'score': float(elems[3]),
                'R': np.array(list(map(float, elems[4].split())), np.float64).reshape((3, 3)),
                't': np.array(list(map(float, elems[5].split())), np.float64).reshape((3, 1)),
                'time': float(elems[6])})
    return results
